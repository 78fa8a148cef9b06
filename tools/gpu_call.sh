cd /root/repo
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err; echo bench rc=$?
python tools/q_time.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_project_q -s 3 -c 1 -f -o gpurun_out/r02_q_final2_full python tools/q_time.py > gpurun_out/ncu_q_final2.log 2>&1; echo ncu rc=$?
python tools/potentials_probe.py > gpurun_out/r02_potentials_probe_e.txt 2>&1; tail -8 gpurun_out/r02_potentials_probe_e.txt
python bench.py --configs > gpurun_out/r02_configs_m.jsonl 2> gpurun_out/r02_configs_m.err; echo configs rc=$?
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_23.txt 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02_pytest_23.txt
