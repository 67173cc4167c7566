mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pic.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log
for v in default pic_single; do
  if [ $v != default ]; then export PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so; fi
  echo "== $v"; for n in 1000000 20000000 100000000; do timeout 300 python tools/pic_profile.py $n 2>&1 | grep "fused step"; done
done
unset PYFDTD_B200_LIB
python - <<'PY'
import sys
sys.path.insert(0, ".")
import torch, bench
import pyfdtd_b200
from pyfdtd_b200 import _native as nat
print(bench.leg_pic(torch, nat, 20_000_000, 6550.0))
PY
