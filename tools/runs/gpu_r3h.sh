mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do
timeout 300 python tools/single_run_profile.py 2>&1 | grep -E "Controller seconds" | cut -c1-120
timeout 300 python tools/single_run_profile.py free 2>&1 | grep -E "Controller seconds" | cut -c1-120
done
