cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29567 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_final2.json 2> gpurun_out/bench_n2_final2.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_n2_final2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus','scaling')}, d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['share_of_step'])
PY
