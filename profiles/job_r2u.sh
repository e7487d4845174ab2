# round-2 (u): batch 2 under sequence parallelism (Wan processors; samples travel one after the other), 2 GPUs
set -x
mkdir -p gpurun_out
for mode in peer nccl; do
VB_ULYSSES=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957${#mode} tests/mgpu_check.py > gpurun_out/r2u_mgpu_$mode.log 2>&1; echo "$mode rc=$?"
grep -v "^W\|^\[W\|Warning" gpurun_out/r2u_mgpu_$mode.log | tail -16
done
