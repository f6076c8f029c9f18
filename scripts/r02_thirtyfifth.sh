# thirty-fifth GPU pass: r = 8 band GEMM for C = 96 / 128 (16-pixel-wide tiles): tests and the configs[4] bench line
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "band_gemm or golden or cfg4 or fp16" 2>&1 | tail -3
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --per-level > $O/r02am_bench_cfg4.json 2> $O/r02am_bench_cfg4.err; tail -6 $O/r02am_bench_cfg4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02am_bench_cfg4.json").read().strip().splitlines()[-1])
print("cfg4", round(d["value"],1), round(d["ms_per_step"],3), d["oracle_check"])
PY
R=8 timeout 120 python scripts/check_fwd_impl.py 2 16 68 120 96 fp16 smooth 2>&1 | sed -E 's/max err.*\); //'
