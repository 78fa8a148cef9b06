cd /root/repo
for v in base elect base elect; do
BC_LIB_PATH=/root/repo/beta-cores_b200/lib/lib_$v.so timeout 300 python tools/q_time.py 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
