mkdir -p gpurun_out
python -m pytest tests/test_gpu_pic.py -m gpu -x -q 2>&1 | tail -3
python tools/pic_profile.py 20000000 | tail -2
for v in pic_lb4 pic_lb4_nounroll pic_nounroll; do echo $v; PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so python tools/pic_profile.py 20000000 | tail -2; done
python __graft_entry__.py smoke 2>&1 | tail -6
