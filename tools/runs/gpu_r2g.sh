mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2g_pytest.log
timeout 300 python tools/nl_profile.py 1024 128 closed 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tools/sweep_multigpu_check.py 2>&1 | grep -v "^\*\|OMP" | tail -6
