# round 2, call b: phase-balanced tile kernel variants (parity subset + headline rate, 512-step passes like bench.py)
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2b_baseline.txt 2>&1
import os, sys, subprocess
for v in ["default", "pb", "pb_c3", "c3"]:
    env = dict(os.environ)
    if v != "default":
        env["PYFDTD_B200_LIB"] = os.path.join(os.getcwd(), "py-fdtd_pic_b200", "variants", f"lib_{v}.so")
    r = subprocess.run([sys.executable, "bench.py", "--no-cpu", "--no-extras", "--steps", "4"], env=env, capture_output=True, text=True, timeout=600)
    import json
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(v, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms_avg"], 4), flush=True)
    except Exception as e:
        print(v, "FAILED", r.stdout[-500:], r.stderr[-1500:], flush=True)
PY
cat gpurun_out/r2b_baseline.txt
for v in pb pb_c3; do
  PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longgrid.py -m gpu -q > gpurun_out/r2b_pytest_$v.log 2>&1; echo "pytest $v rc=$?"
  tail -3 gpurun_out/r2b_pytest_$v.log
done
