python tools/lorentz_profile.py exact | tail -1
for v in t512_k32 t512_k48 t256_k16 t512_k32_c4; do echo $v; PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so timeout 300 python tools/lorentz_profile.py exact | tail -1; done
