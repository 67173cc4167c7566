mkdir -p gpurun_out
for cfg in "0 0" "1 0" "1 1" "0 0" "1 1"; do
  set -- $cfg
  echo "== overlap=$1 nccl_high_prio=$2"
  PF_LONGGRID_OVERLAP=$1 PF_NCCL_HIGH_PRIO=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$1 tools/longgrid_multigpu_check.py --no-check --cells 200000000 --steps 512 2>&1 | grep "rate="
  PF_LONGGRID_OVERLAP=$1 PF_NCCL_HIGH_PRIO=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2964$1 tools/longgrid_multigpu_check.py --no-check --cells 200000000 --steps 512 --mode free 2>&1 | grep "rate="
done
PF_LONGGRID_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 tools/longgrid_multigpu_check.py 2>&1 | grep "single GPU\|rate="
