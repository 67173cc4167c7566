set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r1i_pytest.log
python bench.py > gpurun_out/r1i_bench.json 2> gpurun_out/r1i_bench.err
python bench.py --fp32 --no-extras --no-cpu > gpurun_out/r1i_bench_fp32.json 2>> gpurun_out/r1i_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r1i_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_pic_move$ -s 4 -c 1 -f -o gpurun_out/r1i_pic_move python tools/pic_profile.py 20000000 > gpurun_out/r1i_pic_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_tile$ -s 2 -c 1 -f -o gpurun_out/r1i_lor_fp32 python tools/lorentz_profile.py fp32 > gpurun_out/r1i_lor_fp32_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_tile$ -s 2 -c 1 -f -o gpurun_out/r1i_lor_exact python tools/lorentz_profile.py exact > gpurun_out/r1i_lor_exact_ncu.log 2>&1
tail -3 gpurun_out/r1i_pytest.log; cut -c1-300 gpurun_out/r1i_bench.json; tail -3 gpurun_out/r1i_bench.err
