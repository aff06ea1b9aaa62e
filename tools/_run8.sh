cd /root/repo
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
for i in 1 2; do for f in 0 1; do ACX_PDL=$f python tools/update_time.py 300; done; done
