set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/scale_r1e_n8_balanced.json 2> gpurun_out/scale_r1e_n8_balanced.err; echo "n8 rc=$?"
cut -c1-300 gpurun_out/scale_r1e_n8_balanced.json; tail -3 gpurun_out/scale_r1e_n8_balanced.err
