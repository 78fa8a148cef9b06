cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_8.txt
echo "=== BC_DEBUG library: the whole GPU suite against libbetacores_debug.so (device-side asserts + mbarrier watchdogs) ===" | tee gpurun_out/r02_debug_suite.txt
BC_LIB_PATH=/root/repo/beta-cores_b200/lib/variants/libbetacores_debug.so timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee -a gpurun_out/r02_debug_suite.txt
