mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for p in 0 1 0 1; do
  echo "== PF_TILE_SPLIT=$p"
  PF_TILE_SPLIT=$p timeout 300 python tools/lorentz_profile.py exact 1024 256 2>&1 | tail -1
  PF_TILE_SPLIT=$p timeout 300 python tools/lorentz_profile.py fma 1024 256 2>&1 | tail -1
done
PF_TILE_SPLIT=1 timeout 600 python bench.py --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('bench', d['value'], d['e2e']['value'], r['kernel'], r['kernel_ms_avg'], r['kernel_share_of_step'], r['frac'])"
