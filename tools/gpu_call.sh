cd /root/repo
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r02_bench_n8_e.json 2> gpurun_out/r02_bench_n8_e.err; echo bench8 rc=$?
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n8_e.json').read().strip().splitlines()[-1])
print(json.dumps(d['step_breakdown']['mean_over_ranks'])); print(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['selected_indices'], d['e2e']['value'])
P
