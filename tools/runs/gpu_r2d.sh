# round 2, call d (2 GPUs): full GPU suite incl. the multi-rank tests, bench at N=2 (long-grid legs with NCCL exchange)
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2d_pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2d_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2d_bench_2gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
print("full sweep", {k: d["e2e_full_sweep"][k] for k in ("value", "seconds", "host_setup_s", "build_inputs_s")})
for k, v in (d.get("long_grid") or {}).items():
    print(k, {kk: v[kk] for kk in ("Gcell_updates_per_s", "ms", "kernel_ms_max_over_ranks", "exchange_ms_max_over_ranks", "exchange_share")})
oc = d.get("other_configs") or {}
print({k: v.get("Gcell_updates_per_s", v.get("error")) for k, v in oc.items() if isinstance(v, dict) and k.startswith("nl")}, oc.get("error"), oc.get("trace"))
PY
