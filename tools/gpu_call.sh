cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/q_time.py 2>&1 | tail -1 | tee gpurun_out/r02_q_tab_f.txt
python tools/potentials_probe.py > gpurun_out/r02_potentials_probe_d.txt 2>&1; tail -8 gpurun_out/r02_potentials_probe_d.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_22.txt 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02_pytest_22.txt
