python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
for e in single_run reflection_sweep nonlinear_sweep kerr_lorentz_long_grid pic_beam; do echo "== $e"; timeout 600 python examples/$e.py 2>&1 | tail -3 | cut -c1-200; done
