mkdir -p gpurun_out
timeout 600 python tools/tile_fixed_cost.py 1024 2>&1 | tail -9
timeout 300 python tools/tile_fixed_cost.py 64 2>&1 | tail -2
for v in pic_single; do PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so timeout 300 python tools/pic_profile.py 20000000 2>&1 | grep "fused step"; done
timeout 900 python bench.py --quick-extras --steps 2 2> gpurun_out/r2o_quick.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('quick ok', d['value'], d['config'], list(d['other_configs'].keys()), d['other_configs'].get('error'))
print(d['other_configs'].get('single_run_lorentz_default'))
print(d['other_configs'].get('lorentz_sweep_optional_modes_Gcell_updates_per_s'))"; tail -2 gpurun_out/r2o_quick.err
