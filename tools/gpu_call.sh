cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/q_time.py 2>&1 | tail -1 | tee gpurun_out/r02_q_tab_g.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
