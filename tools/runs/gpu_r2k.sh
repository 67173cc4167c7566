# round 2, call k: full suite, ncu captures of the round (headline k_tile, PIC passes), launch list of a bench run, full bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log
for n in 1000000 20000000 100000000; do timeout 300 python tools/pic_profile.py $n 2>&1 | grep "fused step"; done
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 4 -c 1 -o gpurun_out/r2k_k_tile_lorentz python tools/lorentz_profile.py exact 1024 128 > gpurun_out/r2k_ncu_tile.log 2>&1; tail -1 gpurun_out/r2k_ncu_tile.log
ncu --set full --clock-control none --import-source on -k regex:k_pic_step1 -s 6 -c 2 -o gpurun_out/r2k_k_pic_step1 python tools/pic_profile.py 20000000 > gpurun_out/r2k_ncu_pic.log 2>&1; tail -1 gpurun_out/r2k_ncu_pic.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r2k_ncu_bench.log 2>&1
timeout 1500 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2k_bench.err; cut -c1-300 gpurun_out/r2k_bench.json
