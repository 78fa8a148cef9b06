cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_20.txt 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02_pytest_20.txt
BC_LIB_PATH=$PWD/beta-cores_b200/lib/variants/libbetacores_debug.so python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_debug.txt 2>&1; echo debug pytest rc=$?; tail -4 gpurun_out/r02_pytest_debug.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
