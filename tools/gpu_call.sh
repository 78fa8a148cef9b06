cd /root/repo
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 tools/shard_edge.py 2>&1 | grep "^N \|Error\|error" | head
echo rc=$?
