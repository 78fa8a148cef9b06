cd /root/repo
mkdir -p gpurun_out
for v in default subs2; do
  if [ $v = default ]; then unset BC_LIB_PATH; else export BC_LIB_PATH=/root/repo/beta-cores_b200/lib/variants/libbetacores_$v.so; fi
  timeout 120 python tools/q_time.py 2>&1 | tail -1
done | tee gpurun_out/r02_q2_c.txt
unset BC_LIB_PATH
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_table or tensor_core or precision_tier_contraction or fused_colsum" 2>&1 | tail -4
