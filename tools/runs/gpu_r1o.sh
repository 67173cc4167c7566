mkdir -p gpurun_out
for m in free lorentz lorentz_nl; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_tile$ -s 2 -c 1 -f -o gpurun_out/r1o_long_$m python tools/longgrid_profile.py $m 50000000 > gpurun_out/r1o_long_$m.log 2>&1
  tail -1 gpurun_out/r1o_long_$m.log | cut -c1-200
done
