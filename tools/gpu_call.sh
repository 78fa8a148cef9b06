cd /root/repo
timeout 600 python tools/potentials_probe.py 2>&1 | tail -14
