mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pic.py -m gpu -q > gpurun_out/r2h_pytest_pic.log 2>&1; echo "pytest pic rc=$?"; tail -6 gpurun_out/r2h_pytest_pic.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for n in 1000000 20000000 100000000; do timeout 300 python tools/pic_profile.py $n 2>&1 | tail -2; done
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pic.py -m gpu -q -k "fused_step or coupled_fused" 2>&1 | tail -4
