mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_r01_q_n8.json 2> gpurun_out/bench_r01_q_n8.err; echo "bench8 rc=$?"
python -c "
import json;t=open('gpurun_out/bench_r01_q_n8.json').read();d=json.loads(t[t.index('{'):]);r=d['roofline'];print(d['n_gpus'],d['value'],d['ms_per_step'],r['launch_ms'],r['score_pass_ms'],d['selected_indices'],r['share_of_step'],d['host'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 2 --warmup 3 --no-e2e --rows 100000000 > gpurun_out/bench_r01_q_n8_100M.json 2> gpurun_out/bench_r01_q_n8_100M.err; echo "bench8-100M rc=$?"
python -c "
import json;t=open('gpurun_out/bench_r01_q_n8_100M.json').read();d=json.loads(t[t.index('{'):]);r=d['roofline'];print(d['n_gpus'],d['value'],d['ms_per_step'],r['launch_ms'],r['score_pass_ms'],d['selected_indices'],r['share_of_step'])"
tail -3 gpurun_out/bench_r01_q_n8_100M.err
