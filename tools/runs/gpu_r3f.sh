timeout 300 python tools/single_run_profile.py 2>&1 | head -48 | cut -c1-170
