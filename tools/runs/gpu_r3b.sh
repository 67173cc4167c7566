mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 > gpurun_out/r3b_bench_2gpu.json 2> gpurun_out/r3b_bench_2gpu.err; echo "bench2 rc=$?"; tail -2 gpurun_out/r3b_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3b_bench_2gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "full", round(d["e2e_full_sweep"]["value"], 1), d["e2e_full_sweep"]["seconds"])
for k, v in (d.get("long_grid") or d["other_configs"].get("long_grid") or {}).items():
    print("  ", k, round(v["Gcell_updates_per_s"], 1), "ms", round(v["ms"], 2), "xchg", round(v["exchange_ms_max_over_ranks"], 3))
oc = d["other_configs"]
print({k: round(v["Gcell_updates_per_s"], 1) for k, v in oc.items() if isinstance(v, dict) and "Gcell_updates_per_s" in v}, oc.get("error"))
print({k: round(v["particle_steps_per_s"] / 1e10, 2) for k, v in oc["pic"].items() if isinstance(v, dict) and "particle_steps_per_s" in v})
PY
