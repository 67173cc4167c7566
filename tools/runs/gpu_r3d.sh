mkdir -p gpurun_out
for m in 1 8 32; do echo "== $m members"; timeout 300 python tools/tile_fixed_cost.py $m 2>&1 | tail -8; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv
