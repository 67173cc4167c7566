# round 2, call l: small-tile latency engine for single grids: full suite + single-run profile + step latency
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2l_pytest.log
timeout 600 python tools/single_run_profile.py lorentz > gpurun_out/r2l_single_run_lorentz.txt 2>&1; grep -E "Controller seconds|k_tile:" gpurun_out/r2l_single_run_lorentz.txt; sed -n '/cumulative/,$p' gpurun_out/r2l_single_run_lorentz.txt | head -40
timeout 600 python tools/single_run_profile.py free 2>&1 | grep -E "Controller seconds|k_tile:"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
