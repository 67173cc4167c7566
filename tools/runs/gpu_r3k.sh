export PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_t2048.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for kb in 64 96 128; do echo "k=$kb"; timeout 300 python - <<PY
import sys; sys.path.insert(0, '.')
import torch, bench
b, t = bench.lorentz_sweep_batch(1024, 384, 64)
b.upload(); b.randomize_state(seed=1234)
def step():
    b.reset_state(template=True); b.run(do_pol=True, k_block=$kb)
sec = bench.time_cuda(torch, step, 2)
print({"k": $kb, "Gcell_updates_per_s": b.cell_steps / sec / 1e9})
PY
done
unset PYFDTD_B200_LIB
timeout 300 python tools/lorentz_profile.py exact 1024 384 2>&1 | tail -1
