mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r01_q_n2.json 2> gpurun_out/bench_r01_q_n2.err; echo "bench2 rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_r01_q_n2.json'));r=d['roofline'];print(d['n_gpus'],d['value'],d['ms_per_step'],r['launch_ms'],r['score_pass_ms'],d['e2e']['value'],d['selected_indices'],r['share_of_step'])"
tail -5 gpurun_out/bench_r01_q_n2.err
