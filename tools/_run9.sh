cd /root/repo
for i in 1 2; do for f in 0 1 2; do ACX_PDL=$f python tools/update_time.py 300; done; done
