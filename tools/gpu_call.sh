cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "learn_beta" 2>&1 | tail -15
