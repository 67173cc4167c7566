set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r1f_pytest.log
python bench.py > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err
python bench.py --fp32 --no-extras --no-cpu > gpurun_out/r1f_bench_fp32.json 2>> gpurun_out/r1f_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1f_bench_ref.json 2>> gpurun_out/r1f_bench.err
python tools/fp32_report.py --full --no-speed > gpurun_out/r1f_fp32_accuracy.log 2>&1
tail -3 gpurun_out/r1f_pytest.log; cut -c1-400 gpurun_out/r1f_bench.json; cut -c1-300 gpurun_out/r1f_bench_fp32.json; cut -c1-300 gpurun_out/r1f_bench_ref.json
