for v in default c1 default c1; do
  if [ $v = default ]; then unset PYFDTD_B200_LIB; else export PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so; fi
  echo "== $v"
  timeout 300 python tools/nl_profile.py 1024 128 2>&1 | tail -1 | cut -c1-110
  timeout 300 python tools/nl_profile.py 1024 128 newton 2>&1 | tail -1 | cut -c1-110
  timeout 300 python tools/lorentz_profile.py exact 1024 256 2>&1 | tail -1
done
PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_c1.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
