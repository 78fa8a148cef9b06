cd /root/repo
mkdir -p gpurun_out
for v in default drain default drain; do
  if [ $v = default ]; then unset BC_LIB_PATH; else export BC_LIB_PATH=/root/repo/beta-cores_b200/lib/variants/libbetacores_$v.so; fi
  timeout 120 python tools/q_time.py 2>&1 | tail -1
done | tee gpurun_out/r02_q_variants_c.txt
