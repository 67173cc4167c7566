set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1b_pytest.log
python bench.py > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 4 -c 1 -f -o gpurun_out/r1b_nl_tile python tools/nl_profile.py 256 128 > gpurun_out/r1b_nl_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_pic --csv --log-file gpurun_out/r1b_pic_launches.csv python tools/pic_profile.py 20000000 > gpurun_out/r1b_pic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pic_count|k_pic_move|k_pic_cell_sums" -s 6 -c 3 -f -o gpurun_out/r1b_pic python tools/pic_profile.py 20000000 > gpurun_out/r1b_pic_ncu.log 2>&1
tail -3 gpurun_out/r1b_pytest.log; cat gpurun_out/r1b_bench.json | cut -c1-600
