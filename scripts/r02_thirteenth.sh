# thirteenth GPU pass: evidence for the tensor-workspace backward - full pytest, bench lines, ncu of K1 / K2, launch list
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02n_pytest.log
timeout 400 python bench.py > $O/r02n_bench_default.json 2> $O/r02n_bench_default.err; echo "bench rc=$?"; tail -c 1500 $O/r02n_bench_default.json
timeout 300 python bench.py --mode fwdbwd --steps 20 --warmup 3 --per-level --no-cpu-baseline > $O/r02n_bench_fwdbwd.json 2> $O/r02n_bench_fwdbwd.err; echo "fwdbwd rc=$?"; tail -12 $O/r02n_bench_fwdbwd.err; tail -c 1200 $O/r02n_bench_fwdbwd.json
timeout 300 python bench.py --mode fwdbwd --dtype fp16 --steps 20 --warmup 3 --per-level --no-cpu-baseline --no-e2e > $O/r02n_bench_fwdbwd_fp16.json 2> $O/r02n_bench_fwdbwd_fp16.err; tail -12 $O/r02n_bench_fwdbwd_fp16.err
bash scripts/ncu_capture.sh r02n_bwd_K1_L2 bwd_fused 6 -- env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp32
bash scripts/ncu_capture.sh r02n_bwd_K2_L2 bwd_fused 7 -- env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp32
timeout 200 python bench.py --mode fwdbwd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02n_plain_fwdbwd.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02n_launches_fwdbwd.csv \
    python bench.py --mode fwdbwd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02n_ncu_fwdbwd.log 2>&1
echo "launch list fwdbwd rc=$?"
du -sh $O
