set -e
python scripts/time_level.py 32 3 > gpurun_out/plain_tl.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwd_stream -s 6 -c 1 -f -o gpurun_out/prof_r01_v3a_stream_L2 python scripts/time_level.py 32 3 > gpurun_out/ncu_tl.log 2>&1
tail -2 gpurun_out/plain_tl.log
