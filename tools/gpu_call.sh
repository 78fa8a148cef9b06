cd /root/repo
mkdir -p gpurun_out
./tools/table_eval | tee gpurun_out/r02_table_eval.jsonl
