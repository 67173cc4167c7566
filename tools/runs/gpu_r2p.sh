mkdir -p gpurun_out
echo "== parity, classic kernel (PF_TILE_PERSIST=0)"
PF_TILE_PERSIST=0 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longgrid.py tests/test_gpu_sweep_setup.py tests/test_gpu_pic.py -m gpu -x -q 2>&1 | tail -4
echo "== parity with the persistent kernel forced, 3 CTAs (every CTA walks through many tiles)"
PF_TILE_PERSIST=1 PF_TILE_PERSIST_CTAS=3 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longgrid.py tests/test_gpu_sweep_setup.py -m gpu -x -q 2>&1 | tail -4
echo "== parity with the persistent kernel forced, all CTAs"
PF_TILE_PERSIST=1 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longgrid.py tests/test_gpu_sweep_setup.py -m gpu -x -q 2>&1 | tail -4
for p in 0 1; do
echo "== bench PF_TILE_PERSIST=$p"
PF_TILE_PERSIST=$p timeout 600 python bench.py --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], d['ms_per_step'], r['kernel'], r['kernel_ms_avg'], r['frac'])"
PF_TILE_PERSIST=$p timeout 600 python tools/tile_fixed_cost.py 1024 2>&1 | tail -8
done
