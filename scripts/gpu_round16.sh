timeout 600 python scripts/data_probe.py 2>&1 | tail -10
