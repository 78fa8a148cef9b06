cd /root/repo
mkdir -p gpurun_out
for v in default latefetch r104; do
  if [ $v = default ]; then unset BC_LIB_PATH; else export BC_LIB_PATH=/root/repo/beta-cores_b200/lib/variants/libbetacores_$v.so; fi
  timeout 300 python tools/q_tiers.py 1000000 2>&1 | tail -1 | tee -a gpurun_out/r02_q_tiers.jsonl
done
unset BC_LIB_PATH
timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -40 | tee gpurun_out/r02_pytest_1.txt
