mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r1j_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_pic_move$ -s 10 -c 1 -f -o gpurun_out/r1j_pic_move python tools/pic_profile.py 20000000 > gpurun_out/r1j_pic_ncu.log 2>&1
tail -2 gpurun_out/r1j_ncu_bench.log | cut -c1-200; wc -l gpurun_out/r1j_launches.csv
