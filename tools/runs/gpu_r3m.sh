# round 2, call v (8 GPUs, final code): multi-rank tests (world 4) and bench.py --gpus 8 / 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/r3m_pytest_multirank.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3m_pytest_multirank.log
for n in 4; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n > gpurun_out/r3m_bench_${n}gpu.json 2> gpurun_out/r3m_bench_${n}gpu.err; echo "bench$n rc=$?"; tail -2 gpurun_out/r3m_bench_${n}gpu.err
done
python - <<'PY'
import json
for n in (4,):
    try:
        d = json.loads(open(f"gpurun_out/r3m_bench_{n}gpu.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no line", e); continue
    print(n, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "full", round(d["e2e_full_sweep"]["value"], 1), d["e2e_full_sweep"]["seconds"])
    for k, v in (d.get("long_grid") or {}).items():
        print("  ", k, round(v["Gcell_updates_per_s"], 1), "ms", round(v["ms"], 2), "kernel", round(v["kernel_ms_max_over_ranks"], 2), "xchg", round(v["exchange_ms_max_over_ranks"], 3))
    oc = d.get("other_configs") or {}
    print("  ", {k: round(v["Gcell_updates_per_s"], 1) for k, v in oc.items() if isinstance(v, dict) and "Gcell_updates_per_s" in v}, oc.get("error"))
PY
