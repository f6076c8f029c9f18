# thirty-fourth GPU pass: band GEMM on every 16-bit level with C in {32, 64, 96, 128}
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "band_gemm or fp16 or bf16 or golden or half" 2>&1 | tail -3
timeout 300 python bench.py --dtype fp16 --steps 30 --warmup 5 --per-level --no-cpu-baseline > $O/r02al_bench_fp16.json 2> $O/r02al_bench_fp16.err; tail -6 $O/r02al_bench_fp16.err
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02al_bench_cfg3.json 2>/dev/null
python - <<'PY'
import json
for t in ("fp16","cfg3"):
    d=json.loads(open(f"gpurun_out/r02al_bench_{t}.json").read().strip().splitlines()[-1])
    print(t, round(d["value"],1), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"))
PY
