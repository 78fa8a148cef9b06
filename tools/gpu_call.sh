cd /root/repo
timeout 1000 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --rows 100000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n8_100M_final.json 2> gpurun_out/bench_n8_100M_final.err
echo rc=$?
tail -c 300 gpurun_out/bench_n8_100M_final.err
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_n8_100M_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus','scaling')}, d['host'], d['roofline']['launch_ms'], d['roofline']['share_of_step'])
PY
