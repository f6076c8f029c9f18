# Round profiling recipe (B200_PROFILING.md): plain run first, then (1) launch list of the bench command,
# (2) one --set full capture of the level-2 streaming kernel.  Outputs under gpurun_out/.
set -e
TAG=${1:-r01_v3}
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
python scripts/time_level.py 32 3 > gpurun_out/plain_tl_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwd_stream -s 6 -c 1 -f -o gpurun_out/prof_${TAG}_stream_L2 \
    python scripts/time_level.py 32 3 > gpurun_out/ncu_tl_$TAG.log 2>&1
tail -1 gpurun_out/plain_tl_$TAG.log
