# fourth GPU pass of round 2: the whole evidence set in one call (the earlier passes' gpurun_out was lost with the
# container).  Every line has its own timeout; outputs go to gpurun_out/r02d_*.
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02d_pytest.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3
# driver-shaped default run (configs[2] fp32 forward, e2e + cpu baseline)
timeout 400 python bench.py > $O/r02d_bench_default.json 2> $O/r02d_bench_default.err; echo "bench rc=$?"; tail -c 2500 $O/r02d_bench_default.json; tail -5 $O/r02d_bench_default.err
timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --per-level > $O/r02d_bench_fp32_levels.json 2> $O/r02d_bench_fp32_levels.err; cat $O/r02d_bench_fp32_levels.err
timeout 200 python bench.py --dtype fp16 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --per-level > $O/r02d_bench_fp16_levels.json 2> $O/r02d_bench_fp16_levels.err; cat $O/r02d_bench_fp16_levels.err; tail -c 1200 $O/r02d_bench_fp16_levels.json
timeout 300 python bench.py --mode fwdbwd --steps 10 --warmup 3 --per-level > $O/r02d_bench_fwdbwd.json 2> $O/r02d_bench_fwdbwd.err; echo "fwdbwd rc=$?"; cat $O/r02d_bench_fwdbwd.err | tail -14; tail -c 2500 $O/r02d_bench_fwdbwd.json
timeout 200 python bench.py --mode fwdbwd --dtype fp16 --steps 10 --warmup 3 --per-level --no-e2e --no-cpu-baseline > $O/r02d_bench_fwdbwd_fp16.json 2> $O/r02d_bench_fwdbwd_fp16.err; echo "fwdbwd fp16 rc=$?"; tail -14 $O/r02d_bench_fwdbwd_fp16.err
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --per-level > $O/r02d_bench_cfg4.json 2> $O/r02d_bench_cfg4.err; echo "cfg4 rc=$?"; tail -7 $O/r02d_bench_cfg4.err; tail -c 1500 $O/r02d_bench_cfg4.json
timeout 200 python bench.py --config 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --per-level > $O/r02d_bench_cfg1.json 2> $O/r02d_bench_cfg1.err; echo "cfg1 rc=$?"; tail -14 $O/r02d_bench_cfg1.err
timeout 200 python bench.py --config 0 --steps 50 --warmup 5 --no-cpu-baseline > $O/r02d_bench_cfg0.json 2> $O/r02d_bench_cfg0.err; echo "cfg0 rc=$?"; tail -c 1500 $O/r02d_bench_cfg0.json
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 > $O/r02d_bench_cfg3.json 2> $O/r02d_bench_cfg3.err; echo "cfg3 rc=$?"; tail -c 1500 $O/r02d_bench_cfg3.json; tail -5 $O/r02d_bench_cfg3.err
timeout 200 python bench.py --scaling strong --batch 4 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --per-level > $O/r02d_bench_b4.json 2> $O/r02d_bench_b4.err; echo "b4 rc=$?"; cat $O/r02d_bench_b4.err
timeout 200 python scripts/time_bwd_levels.py 32 fp32 > $O/r02d_time_bwd_fp32.log 2>&1; cat $O/r02d_time_bwd_fp32.log
PWC_BWD_ONLY_FUSED=1 timeout 200 python scripts/time_bwd_levels.py 32 fp16 > $O/r02d_time_bwd_fp16.log 2>&1; cat $O/r02d_time_bwd_fp16.log
# launch lists (plain run first, then under ncu)
timeout 200 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02d_plain_fwd.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02d_launches_fwd.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02d_ncu_fwd.log 2>&1
echo "launch list fwd rc=$?"
timeout 200 python bench.py --mode fwdbwd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02d_plain_fwdbwd.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02d_launches_fwdbwd.csv \
    python bench.py --mode fwdbwd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02d_ncu_fwdbwd.log 2>&1
echo "launch list fwdbwd rc=$?"
# full captures
timeout 100 python scripts/time_level.py 32 3 > $O/r02d_plain_tl.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fwd_stream -s 6 -c 1 -f -o $O/prof_r02d_stream_L2 \
    python scripts/time_level.py 32 3 > $O/r02d_ncu_tl.log 2>&1
echo "ncu stream rc=$?"; cat $O/r02d_plain_tl.log
PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 timeout 120 python scripts/time_bwd_levels.py 32 fp32 > $O/r02d_plain_bwdl2.log 2>&1 && \
PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:bwd_fused -s 6 -c 2 -f -o $O/prof_r02d_bwd_L2 \
    python scripts/time_bwd_levels.py 32 fp32 > $O/r02d_ncu_bwdl2.log 2>&1
echo "ncu bwd rc=$?"
timeout 90 python scripts/check_fwd_impl.py 4 32 112 256 32 fp16 > $O/r02d_plain_mma16.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mma16 -s 3 -c 1 -f -o $O/prof_r02d_mma16_L2 \
    python scripts/check_fwd_impl.py 4 32 112 256 32 fp16 > $O/r02d_ncu_mma16.log 2>&1
echo "ncu mma16 rc=$?"; cat $O/r02d_plain_mma16.log
ls -la $O/*.ncu-rep
