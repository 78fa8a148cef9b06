cd /root/repo
mkdir -p gpurun_out
python tools/q_time.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_project_q -s 3 -c 1 -f -o gpurun_out/r02_q_tab_full python tools/q_time.py > gpurun_out/ncu_q_tab.log 2>&1; echo ncu rc=$?; tail -2 gpurun_out/plain.log
