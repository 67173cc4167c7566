mkdir -p gpurun_out
python bench.py > gpurun_out/r1r_bench.json 2> gpurun_out/r1r_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1r_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r1r_ncu_bench.log 2>&1
tail -2 gpurun_out/r1r_bench.err; cut -c1-200 gpurun_out/r1r_bench.json
