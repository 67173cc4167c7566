# usage: bash tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
{
echo "# gpurun --gpus $N"
$TR --master-port 29501 bench.py --gpus $N --steps 5 --warmup 3 --no-extras 2>&1 | tail -1
$TR --master-port 29502 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -1
for mode in lorentz lorentz_nl; do
  $TR --master-port 29503 tools/longgrid_multigpu_check.py --cells 400000 --steps 256 --mode $mode 2>&1 | grep -E "==|rate="
  $TR --master-port 29504 tools/longgrid_multigpu_check.py --cells $((100000000*N)) --steps 320 --mode $mode --no-check 2>&1 | grep -E "rate="
done
} > gpurun_out/r1_multi_$N.txt 2>&1
cat gpurun_out/r1_multi_$N.txt | cut -c1-700
