set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1d.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_r1d.log
for mode in 1 0; do
VB_ULYSSES_BALANCE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/scale_r1d_n2_balance$mode.json 2> gpurun_out/scale_r1d_n2_balance$mode.err; echo "n2 balance=$mode rc=$?"
cut -c1-400 gpurun_out/scale_r1d_n2_balance$mode.json; tail -3 gpurun_out/scale_r1d_n2_balance$mode.err
done
