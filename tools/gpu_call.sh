cd /root/repo
mkdir -p gpurun_out
BC_LIB_PATH=$PWD/beta-cores_b200/lib/variants/libbetacores_laptrace.so python tools/laplace_parts.py 2>&1 | tail -9 | tee gpurun_out/r02_laplace_parts_b.txt
python -m pytest tests/test_gpu_sampler.py -m gpu -x -q 2>&1 | tail -5
