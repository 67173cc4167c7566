mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/single_run_profile.py 2>&1 | grep -E "Controller seconds|k_tile:|synchronize|prepare_pass|PassRun|finish" | cut -c1-200
timeout 900 python bench.py --quick-extras --steps 2 --no-cpu 2> gpurun_out/r2s_quick.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('quick ok', d['value'], d['other_configs'].get('error'))
print(d['other_configs'].get('single_run_lorentz_default'))
print(d['other_configs'].get('single_run_free_default'))"; tail -2 gpurun_out/r2s_quick.err
