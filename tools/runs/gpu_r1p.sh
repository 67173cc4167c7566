python tools/longgrid_profile.py free 50000000 | tail -1
for v in free_mb4 free_mb5 free_c4_mb4; do echo $v; PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so python tools/longgrid_profile.py free 50000000 | tail -1; done
