mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 300 python tools/single_run_profile.py 2>&1 | grep -E "Controller seconds|k_tile:" | cut -c1-220
timeout 300 python tools/single_run_profile.py free 2>&1 | grep -E "Controller seconds|k_tile:" | cut -c1-220
done
timeout 300 python tools/single_run_profile.py nl 2>&1 | grep -E "Controller seconds|k_tile:" | cut -c1-220
timeout 300 python tools/lorentz_profile.py exact 1024 256 2>&1 | tail -1
