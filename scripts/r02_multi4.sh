# 4-GPU pass: NCCL all-reduce of the training configuration's gradients, the data-parallel training step, strong scaling
set -x
mkdir -p gpurun_out; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 4 --mode allreduce --params sm --steps 50 --warmup 5 > $O/r02q_allreduce_sm_n4.json 2> $O/r02q_allreduce_sm_n4.err; echo "rc=$?"; tail -c 900 $O/r02q_allreduce_sm_n4.json
timeout 300 $TR bench.py --gpus 4 --mode allreduce --params lg --steps 50 --warmup 5 > $O/r02q_allreduce_lg_n4.json 2> $O/r02q_allreduce_lg_n4.err; echo "rc=$?"; tail -c 900 $O/r02q_allreduce_lg_n4.json
timeout 400 $TR bench.py --gpus 4 --config 3 --steps 10 --warmup 3 > $O/r02q_train_n4.json 2> $O/r02q_train_n4.err; echo "rc=$?"; tail -c 1200 $O/r02q_train_n4.json; tail -3 $O/r02q_train_n4.err
timeout 400 $TR bench.py --gpus 4 --scaling strong --steps 50 --warmup 5 --no-cpu-baseline > $O/r02q_strong_n4.json 2> $O/r02q_strong_n4.err; echo "rc=$?"; tail -c 1500 $O/r02q_strong_n4.json
