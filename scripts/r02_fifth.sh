# fifth GPU pass: full pytest (the fourth stopped at the first failure), ncu captures summarised on the box, lane-map A/B
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/r02e_pytest.log
PWC_BWD_ONLY_FUSED=1 timeout 200 python scripts/time_bwd_levels.py 32 fp32 > $O/r02e_time_bwd_fp32.log 2>&1; cat $O/r02e_time_bwd_fp32.log
PWC_B200_LIB=$PWD/tfoptflow_b200/_lib/libpwc_lm0.so PWC_BWD_ONLY_FUSED=1 timeout 200 python scripts/time_bwd_levels.py 32 fp32 2>&1 | sed "s/^/[lanemap0] /"
bash scripts/ncu_capture.sh r02e_bwd_K1_L2 bwd_fused 6 -- env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp32
bash scripts/ncu_capture.sh r02e_bwd_K2_L2 bwd_fused 7 -- env PWC_LEVELS=2 PWC_BWD_ONLY_FUSED=1 python scripts/time_bwd_levels.py 32 fp32
bash scripts/ncu_capture.sh r02e_stream_L2 fwd_stream 6 -- python scripts/time_level.py 32 3
bash scripts/ncu_capture.sh r02e_mma16_L2 mma16 3 -- python scripts/check_fwd_impl.py 4 32 112 256 32 fp16
# launch lists
timeout 200 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02e_plain_fwd.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02e_launches_fwd.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02e_ncu_fwd.log 2>&1
echo "launch list fwd rc=$?"
timeout 200 python bench.py --mode fwdbwd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02e_plain_fwdbwd.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02e_launches_fwdbwd.csv \
    python bench.py --mode fwdbwd --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > $O/r02e_ncu_fwdbwd.log 2>&1
echo "launch list fwdbwd rc=$?"
for k in 0 8 16 32; do
  timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-check --opt side_sms=$k 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('side_sms', '$k', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-check --opt overlap_levels=0 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('overlap 0', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
du -sh $O
