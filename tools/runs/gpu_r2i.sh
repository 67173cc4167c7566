mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pic.py tests/test_gpu_sweep_setup.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2i_pytest.log
for v in default pic_mb4 pic_mb6; do
  if [ $v != default ]; then export PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so; fi
  echo "== $v"; for n in 1000000 20000000 100000000; do timeout 300 python tools/pic_profile.py $n 2>&1 | grep "fused step"; done
done
unset PYFDTD_B200_LIB
python - <<'PY'
import time, sys
sys.path.insert(0, ".")
import numpy as np, torch
import pyfdtd_b200
from pyfdtd_b200 import MasterController as MC
sys.path.insert(0, "tests")
from test_host_layer import build_objects
V, P, C_V, C_P = build_objects(dict(mode="lorentz", freq=6e9, dom=0.7, win=[7000, 8000], source="sine", periods=1.0))
P.Periods = 1.0
for rep in range(3):
    t0 = time.perf_counter()
    MC.LoopedSim(MC.Reporter(), V, P, C_V, C_P, False, 0.7, 7000, 8000, loop=True, Low=6e9, Interval=2e8)
    torch.cuda.synchronize()
    print("LoopedSim 20-point default-geometry sweep: %.4f s" % (time.perf_counter() - t0), MC.LoopedSim.last_sweep[1][:3])
PY
