# final GPU evidence pass of round 2: full pytest, smoke, bench lines of every mode, ncu summaries (K1, K2, streaming
# forward, band GEMM), launch lists of the default and the fwdbwd command
set -x
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/parity_errors.txt
timeout 900 python -m pytest tests -q -m gpu > $O/r02z_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02z_pytest.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee $O/r02z_smoke.log
timeout 400 python bench.py > $O/r02z_bench_default.json 2> $O/r02z_bench_default.err; echo "bench rc=$?"; tail -c 600 $O/r02z_bench_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02z_bench_reference.json 2> $O/r02z_bench_reference.err; echo "ref rc=$?"; tail -c 400 $O/r02z_bench_reference.json
timeout 300 python bench.py --mode fwdbwd --steps 20 --warmup 3 --per-level --no-cpu-baseline > $O/r02z_bench_fwdbwd.json 2> $O/r02z_bench_fwdbwd.err; echo "fwdbwd rc=$?"; tail -12 $O/r02z_bench_fwdbwd.err
timeout 300 python bench.py --dtype fp16 --steps 30 --warmup 5 --per-level --no-cpu-baseline > $O/r02z_bench_fp16.json 2> $O/r02z_bench_fp16.err; echo "fp16 rc=$?"; tail -6 $O/r02z_bench_fp16.err; tail -c 300 $O/r02z_bench_fp16.json
timeout 300 python bench.py --mode fwdbwd --dtype fp16 --steps 20 --warmup 3 --per-level --no-cpu-baseline --no-e2e > $O/r02z_bench_fwdbwd_fp16.json 2> $O/r02z_bench_fwdbwd_fp16.err; tail -12 $O/r02z_bench_fwdbwd_fp16.err
timeout 300 python bench.py --config 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --per-level > $O/r02z_bench_cfg1.json 2> $O/r02z_bench_cfg1.err; tail -13 $O/r02z_bench_cfg1.err
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --per-level > $O/r02z_bench_cfg4.json 2> $O/r02z_bench_cfg4.err; tail -6 $O/r02z_bench_cfg4.err
timeout 300 python bench.py --config 0 --steps 50 --warmup 5 --no-cpu-baseline > $O/r02z_bench_cfg0.json 2> $O/r02z_bench_cfg0.err; tail -c 300 $O/r02z_bench_cfg0.json
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 > $O/r02z_bench_cfg3.json 2> $O/r02z_bench_cfg3.err; tail -c 500 $O/r02z_bench_cfg3.json
bash scripts/ncu_capture.sh r02z_bwd_K1_L2 bwd_fused 6 -- env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp32 > /dev/null
bash scripts/ncu_capture.sh r02z_bwd_K2_L2 bwd_fused 7 -- env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp32 > /dev/null
bash scripts/ncu_capture.sh r02z_stream_L2 fwd_stream 6 -- python scripts/time_level.py 32 3 > /dev/null
timeout 200 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02z_plain_fwd.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02z_launches_fwd.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02z_ncu_fwd.log 2>&1
echo "launch list fwd rc=$?"
timeout 200 python bench.py --mode fwdbwd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02z_plain_fwdbwd.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02z_launches_fwdbwd.csv \
    python bench.py --mode fwdbwd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02z_ncu_fwdbwd.log 2>&1
echo "launch list fwdbwd rc=$?"
du -sh $O
