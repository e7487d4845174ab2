set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/scale_r1o_n8.json 2> gpurun_out/scale_r1o_n8.err; echo "n8 rc=$?"
cut -c1-200 gpurun_out/scale_r1o_n8.json
VB_ULYSSES=peer timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29714 tests/mgpu_check.py > gpurun_out/mgpu_check_r1o_n8.log 2>&1; echo "mgpu rc=$?"
grep -E "OK|MISMATCH|exchange" gpurun_out/mgpu_check_r1o_n8.log
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rescale or dense" 2>&1 | tail -3
