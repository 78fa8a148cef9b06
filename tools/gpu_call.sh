cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "blackbox or groups or hilbert or encoder or learn_beta or fused" 2>&1 | tail -3
python tools/dense_probe.py 1000000 2>&1 | grep colsum
python tools/dense_probe.py 500000 2>&1 | grep colsum
