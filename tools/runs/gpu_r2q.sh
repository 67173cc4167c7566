mkdir -p gpurun_out
for n in 1 2 3 5; do
echo "== fuzz, persistent forced, $n CTAs"
PF_TILE_PERSIST=1 PF_TILE_PERSIST_CTAS=$n timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "random_geometries" --tb=line 2>&1 | grep -E "passed|failed|Error|assert" | cut -c1-600 | head -8
done
PF_TILE_PERSIST=1 ncu --set full --clock-control none --import-source on -k regex:k_tile_p -s 4 -c 1 -o gpurun_out/r2q_k_tile_p python tools/lorentz_profile.py exact 1024 128 > gpurun_out/r2q_ncu_p.log 2>&1; tail -1 gpurun_out/r2q_ncu_p.log
PF_TILE_PERSIST=0 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 4 -c 1 -o gpurun_out/r2q_k_tile python tools/lorentz_profile.py exact 1024 128 > gpurun_out/r2q_ncu_c.log 2>&1; tail -1 gpurun_out/r2q_ncu_c.log
