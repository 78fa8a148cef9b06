cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r02_pytest_12.txt
python tools/c1_profile.py > gpurun_out/r02_c1_profile_g.txt 2>&1; echo rc=$?; head -8 gpurun_out/r02_c1_profile_g.txt
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_c1_launches_d.csv python tools/c1_launches.py 6 > gpurun_out/c1_ncu.log 2>&1; echo ncu rc=$?
timeout 900 python bench.py --configs > gpurun_out/r02_configs_g.jsonl 2> gpurun_out/r02_configs_g.err; echo configs rc=$?
cut -c1-330 gpurun_out/r02_configs_g.jsonl
