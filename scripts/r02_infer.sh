# configs[0]: the whole PWC-Net forward on one 384x512 pair (GPU arm and CPU arm of --mode infer)
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 200 python bench.py --config 0 --mode infer --steps 50 --warmup 5 > $O/r02an_infer_sm.json 2> $O/r02an_infer_sm.err; echo "rc=$?"; tail -c 700 $O/r02an_infer_sm.json; tail -3 $O/r02an_infer_sm.err
timeout 200 python bench.py --config 0 --mode infer --params lg --batch 8 --steps 30 --warmup 5 > $O/r02an_infer_lg_b8.json 2>/dev/null; tail -c 400 $O/r02an_infer_lg_b8.json
timeout 200 python bench.py --config 0 --mode infer --dtype fp16 --batch 8 --steps 30 --warmup 5 > $O/r02an_infer_sm_fp16_b8.json 2>/dev/null; tail -c 400 $O/r02an_infer_sm_fp16_b8.json
timeout 200 python bench.py --impl reference --config 0 --mode infer --steps 3 --warmup 1 > $O/r02an_infer_reference.json 2>/dev/null; tail -c 500 $O/r02an_infer_reference.json
