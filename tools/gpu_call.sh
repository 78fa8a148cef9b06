cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/q_time.py 2>&1 | tail -1 | tee gpurun_out/r02_q_tab_d.txt
python tools/potentials_probe.py > gpurun_out/r02_potentials_probe_c.txt 2>&1; tail -8 gpurun_out/r02_potentials_probe_c.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_table or tensor_core or precision_tier_contraction or fused_colsum" 2>&1 | tail -3
