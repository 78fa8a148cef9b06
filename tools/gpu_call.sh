cd /root/repo
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --no-cpu-baseline --no-e2e --rows 2500000 > gpurun_out/r02_bench_n2_c.json 2> gpurun_out/r02_bench_n2_c.err; echo bench2 rc=$?
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n2_c.json').read().strip().splitlines()[-1])
print(json.dumps(d["step_breakdown"]))
P
