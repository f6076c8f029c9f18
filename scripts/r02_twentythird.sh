# twenty-third GPU pass: band GEMM tests (both tile variants, C = 96, r = 8), fp16 bench with the resident-halo variant
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "band_gemm or fp16 or bf16 or golden or options" 2>&1 | tail -12
timeout 300 python bench.py --dtype fp16 --steps 30 --warmup 5 --per-level --no-cpu-baseline --no-e2e > $O/r02y_bench_fp16.json 2> $O/r02y_bench_fp16.err; tail -6 $O/r02y_bench_fp16.err; tail -c 300 $O/r02y_bench_fp16.json
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02y_bench_cfg3.json 2> $O/r02y_bench_cfg3.err; tail -c 400 $O/r02y_bench_cfg3.json
