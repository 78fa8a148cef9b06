cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "staged_upload or lr_beta_small" 2>&1 | tail -5
