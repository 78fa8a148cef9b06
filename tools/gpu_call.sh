cd /root/repo
for v in notab tab notab tab; do
BC_LIB_PATH=/root/repo/beta-cores_b200/lib/lib_$v.so timeout 300 python tools/q_time.py 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or potentials or project_f" 2>&1 | tail -4
