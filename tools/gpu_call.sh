cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_27.txt 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02_pytest_27.txt
