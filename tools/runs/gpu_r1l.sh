python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline or heterogeneous or full_size" 2>&1 | tail -5
python bench.py --no-extras > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; tail -3 gpurun_out/r1l_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r1l_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'])
P
