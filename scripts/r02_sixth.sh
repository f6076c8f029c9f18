# sixth GPU pass (session 3): full pytest after the persistent-backward commits, backward time split by skipping one
# piece at a time (results wrong, timing only), forward traces of the level-3 / level-4 shapes, default bench
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02f_pytest.log
PWC_BWD_ONLY_FUSED=1 timeout 200 python scripts/time_bwd_levels.py 32 fp32 2>&1 | tee $O/r02f_time_bwd_fp32.log
for d in 1 2 4 8 16 6 7 14 22 31; do
  PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 timeout 100 python scripts/time_bwd_levels.py 32 fp32 bwd_debug=$d 2>&1 | grep "level" | sed "s/^/[dbg $d] /"
done 2>&1 | tee $O/r02f_bwd_dbg.log
for shp in 56,128,64 28,64,96 112,256,32; do
  SHAPE=$shp T0=1000000000 timeout 100 python scripts/trace_stream.py 32 2>&1 | sed "s/^/[trace $shp] /"
done > $O/r02f_traces.log 2>&1
tail -60 $O/r02f_traces.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --per-level > $O/r02f_bench_fwd.json 2> $O/r02f_bench_fwd.err; cat $O/r02f_bench_fwd.err | tail -8
du -sh $O
