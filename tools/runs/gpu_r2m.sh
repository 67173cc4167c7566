mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2m_pytest.log
timeout 600 python tools/single_run_profile.py lorentz 2>&1 | grep -E "Controller seconds|k_tile:"
timeout 1500 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2m_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], {k: d["roofline"][k] for k in ("bound", "frac", "traffic", "kernel_share_of_step")})
oc = d["other_configs"]
print({k: oc[k]["seconds_e2e"] for k in ("single_run_free_default", "single_run_lorentz_default")})
print({k: (oc[k]["Gcell_updates_per_s"], oc[k]["rank0_roofline"]["fp64"]["frac"]) for k in ("nl_cubic_sweep_closed_form", "nl_cubic_sweep_newton")})
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r2m_bench_reference.json
