# A/B on one box: HEAD's k_tile | tile classes from the table (PML profiles in registers) | ... (PML profiles of the slab+CPML body in smem)
mkdir -p gpurun_out
echo "== parity (default lib)"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longgrid.py tests/test_gpu_sweep_setup.py tests/test_gpu_pic.py -m gpu -x -q 2>&1 | tail -3
for v in head default head default; do
  if [ $v = default ]; then unset PYFDTD_B200_LIB; else export PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so; fi
  echo "== $v"
  timeout 600 python bench.py --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('bench', d['value'], d['ms_per_step'], r['kernel'], r['kernel_ms_avg'], r['frac'])"
  timeout 600 python tools/tile_fixed_cost.py 1024 2>&1 | tail -1
done
unset PYFDTD_B200_LIB
for v in head default; do
  if [ $v = default ]; then unset PYFDTD_B200_LIB; else export PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so; fi
  echo "== $v: other workloads"
  timeout 300 python tools/nl_profile.py 2>&1 | tail -2
  timeout 300 python tools/longgrid_profile.py 2>&1 | tail -3
  timeout 300 python tools/single_run_profile.py 2>&1 | tail -3
done
