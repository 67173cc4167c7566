# round 2, call a: default-lib GPU tests (ABI v2) + warp-exchange kernel variants (parity subset + headline rate)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_default.log 2>&1; echo "pytest default rc=$?" | tee -a gpurun_out/r2a_summary.txt
timeout 300 python tools/lorentz_profile.py exact 1024 2>&1 | tail -1 | tee -a gpurun_out/r2a_summary.txt
for v in wx wx_nsw wx_c4 wx_c4_t2048 wx_c2_t2048; do
  echo "== $v" | tee -a gpurun_out/r2a_summary.txt
  PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so timeout 300 python tools/lorentz_profile.py exact 1024 2>&1 | tail -1 | tee -a gpurun_out/r2a_summary.txt
  PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longgrid.py -m gpu -q > gpurun_out/r2a_pytest_$v.log 2>&1; echo "pytest $v rc=$?" | tee -a gpurun_out/r2a_summary.txt
  tail -3 gpurun_out/r2a_pytest_$v.log | tee -a gpurun_out/r2a_summary.txt
done
