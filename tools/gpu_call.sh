cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "snnls or hilbert or blackbox or nan_and_tie" 2>&1 | tail -3
python tools/dense_probe.py 500000 2>&1 | tail -6
python tools/dense_probe.py 1000000 2>&1 | tail -6
