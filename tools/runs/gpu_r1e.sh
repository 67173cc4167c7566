set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "newton or kerr or fp32 or fma" 2>&1 | tail -15 > gpurun_out/r1e_pytest.log
python tools/fp32_report.py --no-acc --only=fp64_newton,fp32 2>&1 | tail -1 > gpurun_out/r1e_speed_main.log
for v in newton_noinline f32_c2_b4 f32_c2_b3 f32_c1_b2 f32_c4_b3; do
  PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so python tools/fp32_report.py --no-acc --only=fp64_newton,fp32 2>&1 | tail -1 > gpurun_out/r1e_speed_$v.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_tile$ -s 2 -c 1 -f -o gpurun_out/r1e_nl_newton python tools/nl_profile.py 256 128 newton > gpurun_out/r1e_nl_newton_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_tile$ -s 2 -c 1 -f -o gpurun_out/r1e_lor_fp32 python tools/lorentz_profile.py fp32 > gpurun_out/r1e_lor_fp32_ncu.log 2>&1
tail -5 gpurun_out/r1e_pytest.log; cat gpurun_out/r1e_speed_*.log | cut -c1-300
