mkdir -p gpurun_out
for v in DEC DECFEW; do
  BC_LIB_PATH=$PWD/beta-cores_b200/lib/libbc_$v.so timeout 120 python tools/q_time.py
done 2>&1 | grep "ms/pass"
