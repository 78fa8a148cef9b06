cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_n4_final.json 2> gpurun_out/bench_n4_final.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_n4_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus','scaling')}, d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['share_of_step'], d['selected_indices'])
PY
