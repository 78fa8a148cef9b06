cd /root/repo
timeout 60 ./tools/mma_i8_rate
