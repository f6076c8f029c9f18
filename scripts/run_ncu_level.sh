# usage: run_ncu_level.sh <out-name> [time_level args...]   (plain run first, then one ncu --set full capture)
set -e
out=$1; shift
python scripts/time_level.py 32 3 "$@" > gpurun_out/plain_tl.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwd_stream -s 6 -c 1 -f -o gpurun_out/$out python scripts/time_level.py 32 3 "$@" > gpurun_out/ncu_tl.log 2>&1
tail -1 gpurun_out/plain_tl.log
