set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fp32 or fma" 2>&1 | tail -15 > gpurun_out/r1c_pytest.log
python tools/fp32_report.py --full > gpurun_out/r1c_fp32_report.log 2>&1
for v in f32_c4_b4 f32_c2_b4 f32_c8_b2 f32_c2_b2; do
  PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so python tools/fp32_report.py --no-acc 2>&1 | tail -1 > gpurun_out/r1c_speed_$v.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tile<" -s 2 -c 1 -f -o gpurun_out/r1c_nl_tile python tools/nl_profile.py 256 128 > gpurun_out/r1c_nl_ncu.log 2>&1
tail -5 gpurun_out/r1c_pytest.log; tail -4 gpurun_out/r1c_fp32_report.log; cat gpurun_out/r1c_speed_*.log
