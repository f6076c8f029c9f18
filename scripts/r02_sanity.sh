# sanity pass at HEAD: full GPU test suite, smoke, the driver's default bench command, ncu summary of the band GEMM
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/r02zz_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02zz_pytest.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py > $O/r02zz_bench_default.json 2> $O/r02zz_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zz_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
PY
bash scripts/ncu_capture.sh r02ai_band_L2 mma16_band 3 -- python scripts/check_fwd_impl.py 4 32 112 256 32 fp16 > /dev/null
sed -n 1,12p $O/r02ai_band_L2.summary.txt; tail -8 $O/r02ai_band_L2.summary.txt
