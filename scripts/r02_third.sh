# third GPU pass: fused backward with L2 prefetch + warp split variants, launch list, ncu captures, r=8
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
timeout 60 python scripts/check_bwd_one.py 2 30 64 32 stress || exit 1
PWC_BWD_ONLY_FUSED=1 timeout 120 python scripts/time_bwd_levels.py 32 fp32 > gpurun_out/r02c_time_bwd_fp32.log 2>&1; cat gpurun_out/r02c_time_bwd_fp32.log
for v in a44 a86; do
  PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_$v.so PWC_LEVELS=3,2 PWC_BWD_ONLY_FUSED=1 timeout 120 python scripts/time_bwd_levels.py 32 fp32 2>&1 | sed "s/^/[$v] /"
done
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/r02c_pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -8 gpurun_out/r02c_pytest_all.log
timeout 200 python bench.py --config 4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --per-level > gpurun_out/r02c_bench_cfg4.log 2>gpurun_out/r02c_bench_cfg4.err; echo "bench cfg4 rc=$?"; tail -c 1200 gpurun_out/r02c_bench_cfg4.log; tail -7 gpurun_out/r02c_bench_cfg4.err
# launch list of the fwd+bwd step
timeout 200 python bench.py --mode fwdbwd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > gpurun_out/r02c_plain_fwdbwd.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches_fwdbwd.csv \
    python bench.py --mode fwdbwd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > gpurun_out/r02c_ncu_fwdbwd.log 2>&1
echo "launch list rc=$?"
# full captures: K1 + K2 at level 2, the fp16 tensor-core forward
PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 timeout 120 python scripts/time_bwd_levels.py 32 fp32 > gpurun_out/r02c_plain_bwdl2.log 2>&1 && \
PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:bwd_fused -s 6 -c 2 -f -o gpurun_out/prof_r02c_bwd_L2 \
    python scripts/time_bwd_levels.py 32 fp32 > gpurun_out/r02c_ncu_bwdl2.log 2>&1
echo "ncu bwd rc=$?"
timeout 90 python scripts/check_fwd_impl.py 4 32 112 256 32 fp16 > gpurun_out/r02c_plain_mma16.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mma16 -s 3 -c 1 -f -o gpurun_out/prof_r02c_mma16_L2 \
    python scripts/check_fwd_impl.py 4 32 112 256 32 fp16 > gpurun_out/r02c_ncu_mma16.log 2>&1
echo "ncu mma16 rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
