# round 2, call c: full GPU test suite (new sweep-setup tests), quick bench with every leg, full bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest.log
timeout 900 python bench.py --quick-extras --steps 2 > gpurun_out/r2c_bench_quick.json 2> gpurun_out/r2c_bench_quick.err; echo "quick bench rc=$?"; tail -3 gpurun_out/r2c_bench_quick.err; cut -c1-600 gpurun_out/r2c_bench_quick.json
timeout 1500 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", {k: d["roofline"][k] for k in ("bound", "achieved", "peak", "frac", "kernel_ms_avg")})
print("full sweep", d.get("e2e_full_sweep"))
print("cpu", d.get("cpu_baseline"))
for k, v in (d.get("long_grid") or {}).items():
    print(k, {kk: v[kk] for kk in ("Gcell_updates_per_s", "ms", "kernel_ms_max_over_ranks", "exchange_ms_max_over_ranks")}, v["rank0_roofline"]["fp64"]["frac"] if v.get("rank0_roofline") else None)
oc = d.get("other_configs") or {}
print({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk in ("Gcell_updates_per_s", "error", "trace", "seconds_e2e")}) for k, v in oc.items() if k != "pic"})
print("pic", {k: (v["particle_steps_per_s"], v["hbm"]["frac"]) for k, v in (oc.get("pic") or {}).items()})
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2c_bench_reference.json
