# thirty-second GPU pass: bench lines of the 16-bit configurations with the packed-corner gather
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests -q -m gpu -k "band_gemm or fp16 or golden or cfg4" 2>&1 | tail -3
timeout 300 python bench.py --dtype fp16 --steps 30 --warmup 5 --per-level --no-cpu-baseline > $O/r02ah_bench_fp16.json 2> $O/r02ah_bench_fp16.err; tail -6 $O/r02ah_bench_fp16.err
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --per-level > $O/r02ah_bench_cfg4.json 2> $O/r02ah_bench_cfg4.err; tail -6 $O/r02ah_bench_cfg4.err
timeout 300 python bench.py --mode fwdbwd --dtype fp16 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02ah_bench_fwdbwd_fp16.json 2>/dev/null
python - <<'PY'
import json
for t in ("fp16","cfg4","fwdbwd_fp16"):
    d=json.loads(open(f"gpurun_out/r02ah_bench_{t}.json").read().strip().splitlines()[-1])
    print(t, round(d["value"],1), round(d["ms_per_step"],3), round(d["roofline"]["frac"],3), round(d["roofline"]["kernel_ms"]*1e3,1))
PY
