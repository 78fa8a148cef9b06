cd /root/repo
mkdir -p gpurun_out
python tools/c1_profile.py > gpurun_out/r02_c1_profile_i.txt 2>&1; echo rc=$?; head -8 gpurun_out/r02_c1_profile_i.txt
nproc; uptime
