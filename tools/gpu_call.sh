cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r02_pytest_14.txt
timeout 900 python bench.py --configs > gpurun_out/r02_configs_i.jsonl 2> gpurun_out/r02_configs_i.err; echo configs rc=$?
cut -c1-330 gpurun_out/r02_configs_i.jsonl
