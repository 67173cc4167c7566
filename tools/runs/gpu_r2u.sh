# round 2, call u: state after the class-table / two-pass-overlap commits -- full suite, sanitizer on the fuzz tests,
# ncu captures of the headline kernel, launch list of a bench run, full bench (timed), reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "random_geometries or independent_of_time_blocking" > gpurun_out/r2u_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2u_sanitizer_memcheck.log
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 5 -c 1 -o gpurun_out/r2u_k_tile_lorentz python tools/lorentz_profile.py exact 1024 128 > gpurun_out/r2u_ncu_tile.log 2>&1; tail -1 gpurun_out/r2u_ncu_tile.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r2u_ncu_bench.log 2>&1
( time timeout 1500 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 gpurun_out/r2u_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2u_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], {k: d["roofline"][k] for k in ("bound", "frac", "traffic", "kernel_share_of_step")})
print("full sweep", d["e2e_full_sweep"]["value"], d["e2e_full_sweep"]["seconds"], "cpu", d["cpu_baseline"]["value"])
oc = d["other_configs"]
print({k: oc[k]["seconds_e2e"] for k in ("single_run_free_default", "single_run_lorentz_default")})
print({k: (oc[k]["Gcell_updates_per_s"], oc[k]["rank0_roofline"]["fp64"]["frac"]) for k in ("nl_cubic_sweep_closed_form", "nl_cubic_sweep_newton")})
print({k: v for k, v in oc["long_grid"].items() if not isinstance(v, dict)} if isinstance(oc.get("long_grid"), dict) else list(oc.keys()))
PY
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2u_bench_reference.json 2>/dev/null ) 2>&1 | grep real; cut -c1-200 gpurun_out/r2u_bench_reference.json
