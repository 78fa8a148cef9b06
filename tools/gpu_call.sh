cd /root/repo
for v in b40 b21 b20 b40 b21 b20; do
BC_LIB_PATH=/root/repo/beta-cores_b200/lib/lib_$v.so timeout 120 python tools/q_time.py 2>&1 | tail -1
done
BC_LIB_PATH=/root/repo/beta-cores_b200/lib/lib_b21.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or tensor_core or score_nan" 2>&1 | tail -3
