mkdir -p gpurun_out
timeout 900 python tools/class_cost.py 1024 256 2>&1 | tail -8
