cd /root/repo
mkdir -p gpurun_out
python tools/c1_profile.py > gpurun_out/r02_c1_profile_d.txt 2>&1; echo rc=$?; head -40 gpurun_out/r02_c1_profile_d.txt
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -q -x 2>&1 | tail -5
