mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sweep_setup.py -m gpu -x -q 2>&1 | tail -2
for p in 0 1 0 1; do
  echo "== PF_TILE_ORDER=$p"
  PF_TILE_ORDER=$p timeout 300 python tools/lorentz_profile.py exact 1024 512 2>&1 | tail -1
  PF_TILE_ORDER=$p timeout 300 python tools/nl_profile.py 2>&1 | tail -1
done
