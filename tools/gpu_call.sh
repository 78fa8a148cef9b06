cd /root/repo
timeout 300 python tools/sampler_prof.py 2>&1 | grep -v "^$" | head -60
