# usage: bash tools/gpu_multi2.sh N   (under gpurun --gpus N): all BASELINE configs at N GPUs
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"; fi
$RUN tools/bench_configs.py 2> gpurun_out/r1_configs_$N.err | grep '^{' > gpurun_out/r1_configs_$N.jsonl
tail -3 gpurun_out/r1_configs_$N.err; cut -c1-260 gpurun_out/r1_configs_$N.jsonl
