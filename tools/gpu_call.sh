cd /root/repo
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r02_bench_n8_a.json 2> gpurun_out/r02_bench_n8_a.err; echo bench8 rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_n4_a.json 2> gpurun_out/r02_bench_n4_a.err; echo bench4 rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --sweep default --sweep-out gpurun_out/r02_sweep_n8_b.jsonl > gpurun_out/r02_sweep_n8_b.log 2> gpurun_out/r02_sweep_n8_b.err; echo sweep8 rc=$?
