cd /root/repo
timeout 600 python -m pytest tests/test_aux_rows.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python scratch/attr_time.py 2>&1 | tail -2
