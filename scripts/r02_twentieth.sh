# twentieth GPU pass: 16-bit band GEMM for search_range 8 (configs[4])
set -x
mkdir -p gpurun_out; O=gpurun_out
R=8 timeout 120 python scripts/check_fwd_impl.py 4 2 34 60 32 fp16 stress
R=8 timeout 120 python scripts/check_fwd_impl.py 2 2 34 60 32 fp16 stress
R=8 timeout 120 python scripts/check_fwd_impl.py 4 1 17 30 32 fp16 none
R=8 timeout 120 python scripts/check_fwd_impl.py 4 16 272 480 32 fp16 smooth
R=8 timeout 120 python scripts/check_fwd_impl.py 2 16 272 480 32 fp16 smooth
timeout 600 python -m pytest tests -q -m gpu -k "golden or cfg4 or r8 or search_range or fp16 or ranges" 2>&1 | tail -4
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --per-level > $O/r02v_bench_cfg4.json 2> $O/r02v_bench_cfg4.err; tail -6 $O/r02v_bench_cfg4.err; tail -c 400 $O/r02v_bench_cfg4.json
