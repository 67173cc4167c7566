mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m pytest tests/test_gpu_longgrid.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -4
for ov in 0 1 0 1; do
  echo "== overlap=$ov"
  PF_LONGGRID_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2960$ov tools/longgrid_multigpu_check.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -3
  PF_LONGGRID_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$ov tools/longgrid_multigpu_check.py --no-check --cells 200000000 --steps 512 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -2
  PF_LONGGRID_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$ov tools/longgrid_multigpu_check.py --no-check --cells 200000000 --steps 512 --mode free 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -2
done
