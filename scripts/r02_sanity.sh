# sanity pass at HEAD: full GPU test suite, smoke, the driver's default bench command
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/r02zz_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02zz_pytest.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py > $O/r02zz_bench_default.json 2> $O/r02zz_bench_default.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/r02zz_bench_default.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
