# second GPU pass of round 2 (tight timeouts: nothing here may run longer than its line says)
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
for args in "2 9 33 32 stress" "1 37 70 64 stress" "1 20 36 32 smooth" "4 28 64 96 smooth"; do
  timeout 60 python scripts/check_bwd_one.py $args || { echo "check_bwd_one $args FAILED rc=$?"; exit 1; }
done
PWC_BWD_ONLY_FUSED=1 timeout 120 python scripts/time_bwd_levels.py 32 fp32 > gpurun_out/r02b_time_bwd_fp32.log 2>&1; echo "time rc=$?"; cat gpurun_out/r02b_time_bwd_fp32.log
timeout 60 python scripts/time_level.py 32 3 > gpurun_out/r02b_time_l2.log 2>&1; cat gpurun_out/r02b_time_l2.log
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/r02b_pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -15 gpurun_out/r02b_pytest_all.log
for args in "4 8 112 256 32 fp16" "4 8 56 128 64 fp16" "4 2 37 70 32 fp16 stress" "4 1 9 33 64 fp16 none"; do
  timeout 60 python scripts/check_fwd_impl.py $args 2>&1 | tail -3
done
timeout 90 python scripts/check_fwd_impl.py 4 32 112 256 32 fp16 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_default.log 2>gpurun_out/r02b_bench_default.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r02b_bench_default.log; tail -5 gpurun_out/r02b_bench_default.err
for k in 0 8 24; do
  timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-check --opt side_sms=$k 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('side_sms', '$k', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
timeout 200 python bench.py --mode fwdbwd --steps 10 --warmup 3 --no-cpu-baseline --per-level > gpurun_out/r02b_bench_fwdbwd.log 2>gpurun_out/r02b_bench_fwdbwd.err; echo "bench fwdbwd rc=$?"; tail -c 2500 gpurun_out/r02b_bench_fwdbwd.log; tail -12 gpurun_out/r02b_bench_fwdbwd.err
timeout 120 python bench.py --dtype fp16 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --per-level > gpurun_out/r02b_bench_fp16.log 2>gpurun_out/r02b_bench_fp16.err; echo "bench fp16 rc=$?"; tail -c 1500 gpurun_out/r02b_bench_fp16.log; tail -6 gpurun_out/r02b_bench_fp16.err
timeout 200 python bench.py --config 4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --per-level > gpurun_out/r02b_bench_cfg4.log 2>gpurun_out/r02b_bench_cfg4.err; echo "bench cfg4 rc=$?"; tail -c 1500 gpurun_out/r02b_bench_cfg4.log; tail -6 gpurun_out/r02b_bench_cfg4.err
timeout 200 python bench.py --config 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --per-level > gpurun_out/r02b_bench_cfg1.log 2>gpurun_out/r02b_bench_cfg1.err; echo "bench cfg1 rc=$?"; tail -c 1500 gpurun_out/r02b_bench_cfg1.log; tail -14 gpurun_out/r02b_bench_cfg1.err
