mkdir -p gpurun_out
for v in default cubic1 newton1 default cubic1 newton1; do
  if [ $v = default ]; then unset PYFDTD_B200_LIB; else export PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so; fi
  echo "== $v"
  timeout 300 python tools/nl_profile.py 1024 128 2>&1 | tail -1 | cut -c1-120
  timeout 300 python tools/nl_profile.py 1024 128 newton 2>&1 | tail -1 | cut -c1-120
done
