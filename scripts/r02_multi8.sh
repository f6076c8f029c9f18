# 8-GPU pass: NCCL all-reduce of the training configuration's gradients, the data-parallel training step, strong scaling
set -x
mkdir -p gpurun_out; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 8 --mode allreduce --params sm --steps 50 --warmup 5 > $O/r02q_allreduce_sm_n8.json 2> $O/r02q_allreduce_sm_n8.err; echo "rc=$?"; tail -c 900 $O/r02q_allreduce_sm_n8.json
timeout 300 $TR bench.py --gpus 8 --mode allreduce --params lg --steps 50 --warmup 5 > $O/r02q_allreduce_lg_n8.json 2> $O/r02q_allreduce_lg_n8.err; echo "rc=$?"; tail -c 900 $O/r02q_allreduce_lg_n8.json
timeout 400 $TR bench.py --gpus 8 --config 3 --steps 10 --warmup 3 > $O/r02q_train_n8.json 2> $O/r02q_train_n8.err; echo "rc=$?"; tail -c 1200 $O/r02q_train_n8.json; tail -3 $O/r02q_train_n8.err
timeout 400 $TR bench.py --gpus 8 --scaling strong --steps 50 --warmup 5 --no-cpu-baseline > $O/r02q_strong_n8.json 2> $O/r02q_strong_n8.err; echo "rc=$?"; tail -c 1500 $O/r02q_strong_n8.json
timeout 400 $TR bench.py --gpus 8 --mode fwdbwd --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02q_fwdbwd_n8.json 2> $O/r02q_fwdbwd_n8.err; echo "rc=$?"; tail -c 700 $O/r02q_fwdbwd_n8.json
