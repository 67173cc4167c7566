mkdir -p gpurun_out
for v in default picplace5 picplace6 default picplace5 picplace6; do
  if [ $v = default ]; then unset PYFDTD_B200_LIB; else export PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so; fi
  echo "== $v"
  for n in 1000000 20000000; do timeout 300 python tools/pic_profile.py $n 2>&1 | grep -E "fused step|k_pic_step1" | head -4; done
done
