# seventeenth GPU pass: tall levels on streams of their own (overlap_levels 2), side_sms sweep, 16-bit band GEMM on level 3
set -x
mkdir -p gpurun_out; O=gpurun_out
for ov in 1 2; do for k in 0 8 16 24; do
  timeout 120 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --opt overlap_levels=$ov --opt side_sms=$k 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('overlap', $ov, 'side_sms', $k, round(d['value']), round(d['ms_per_step']*1e3,1), 'us', d['oracle_check']['ok'])"
done; done | tee $O/r02s_overlap.log
timeout 90 python scripts/check_fwd_impl.py 4 32 56 128 64 fp16
timeout 90 python scripts/check_fwd_impl.py 3 32 56 128 64 fp16
timeout 90 python scripts/check_fwd_impl.py 4 32 112 256 32 fp16
timeout 90 python scripts/check_fwd_impl.py 3 32 112 256 32 fp16
