# twenty-first GPU pass: r = 8 band GEMM for C = 64, pixel threshold
set -x
mkdir -p gpurun_out; O=gpurun_out
R=8 timeout 120 python scripts/check_fwd_impl.py 4 2 34 60 64 fp16 stress
R=8 timeout 120 python scripts/check_fwd_impl.py 4 1 17 30 64 fp16 none
R=8 timeout 120 python scripts/check_fwd_impl.py 4 16 136 240 64 fp16 smooth
R=8 timeout 120 python scripts/check_fwd_impl.py 2 16 136 240 64 fp16 smooth
timeout 600 python -m pytest tests -q -m gpu -k "golden or cfg4 or r8 or search_range or fp16 or ranges" 2>&1 | tail -4
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --per-level > $O/r02w_bench_cfg4.json 2> $O/r02w_bench_cfg4.err; tail -6 $O/r02w_bench_cfg4.err; tail -c 300 $O/r02w_bench_cfg4.json
