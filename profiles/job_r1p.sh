set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r1p.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r1p.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29715 bench.py --gpus 2 --workload hunyuan --steps 2 --warmup 3 > gpurun_out/scale_r1p_hunyuan_n2.json 2> gpurun_out/scale_r1p_hunyuan_n2.err; echo "hunyuan n2 rc=$?"
cut -c1-300 gpurun_out/scale_r1p_hunyuan_n2.json; tail -3 gpurun_out/scale_r1p_hunyuan_n2.err
