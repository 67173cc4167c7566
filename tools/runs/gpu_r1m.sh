mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r1m_pytest.log
python bench.py > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1m_bench_ref.json 2>> gpurun_out/r1m_bench.err
python __graft_entry__.py smoke 2>&1 | tail -5 > gpurun_out/r1m_smoke.log
cat gpurun_out/r1m_pytest.log gpurun_out/r1m_smoke.log; tail -3 gpurun_out/r1m_bench.err
