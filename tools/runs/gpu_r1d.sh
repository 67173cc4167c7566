set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1d_pytest.log
python tools/fp32_report.py --full > gpurun_out/r1d_fp32_report.log 2>&1
for v in f32_c2_b4 f32_c2_b3 f32_c1_b2 f32_c4_b3; do
  PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so python tools/fp32_report.py --no-acc --only=fp32 2>&1 | tail -1 > gpurun_out/r1d_speed_$v.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_tile$ -s 2 -c 1 -f -o gpurun_out/r1d_nl_tile python tools/nl_profile.py 256 128 > gpurun_out/r1d_nl_ncu.log 2>&1
tail -5 gpurun_out/r1d_pytest.log; tail -6 gpurun_out/r1d_fp32_report.log | cut -c1-300; cat gpurun_out/r1d_speed_*.log
