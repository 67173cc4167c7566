# round 2, call e: inlined fast closed-form cubic law (parity + rate vs the out-of-line round-1 arrangement), PIC-in-medium tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2e_pytest.log
for v in default nl_ool; do
  if [ $v != default ]; then export PYFDTD_B200_LIB=$PWD/py-fdtd_pic_b200/variants/lib_$v.so; fi
  echo "== $v"; timeout 300 python tools/nl_profile.py 1024 128 closed 2>&1 | tail -1
done
unset PYFDTD_B200_LIB
timeout 300 python tools/nl_profile.py 1024 128 newton 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 2 -c 1 -o gpurun_out/r2e_nl_closed python tools/nl_profile.py 256 64 closed > gpurun_out/r2e_ncu_nl.log 2>&1; tail -2 gpurun_out/r2e_ncu_nl.log
