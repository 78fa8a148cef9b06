cd /root/repo
mkdir -p gpurun_out
python tools/q_time.py 2>&1 | tail -1 | tee gpurun_out/r02_q_tab_c.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_19.txt 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r02_pytest_19.txt
python tools/laplace_parts.py 2>&1 | tail -4
python tools/q_time.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_project_q -s 3 -c 1 -f -o gpurun_out/r02_q_final_full python tools/q_time.py > gpurun_out/ncu_q_final.log 2>&1; echo ncu rc=$?
