# round 2, final state: full suite, bench (timed), reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2final_pytest.log
( time timeout 1500 python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err ) 2>&1 | grep real; tail -2 gpurun_out/r2final_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2final_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], {k: d["roofline"][k] for k in ("bound", "frac", "traffic", "kernel_share_of_step")}, d["clocks"])
print("full sweep", d["e2e_full_sweep"]["value"], d["e2e_full_sweep"]["seconds"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
oc = d["other_configs"]
print({k: oc[k]["seconds_e2e"] for k in ("single_run_free_default", "single_run_lorentz_default")})
print({k: (oc[k]["Gcell_updates_per_s"], oc[k]["rank0_roofline"]["fp64"]["frac"]) for k in ("nl_cubic_sweep_closed_form", "nl_cubic_sweep_newton")})
print("pic", json.dumps(oc["pic"])[:600])
for k, v in (d.get("long_grid") or {}).items():
    print("  ", k, round(v["Gcell_updates_per_s"], 1), "ms", round(v["ms"], 2))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2final_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r2final_bench_reference.json
