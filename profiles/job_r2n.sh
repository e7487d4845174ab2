# round-2 (n): 4 GPUs — exchange modes bit-identical to 1 GPU, Wan-14B bench line (the driver's SCALE run includes N = 4)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "4-" 2>&1 | tail -40 > gpurun_out/r2n_pytest_4gpu.log; tail -8 gpurun_out/r2n_pytest_4gpu.log
grep -q "failed" gpurun_out/r2n_pytest_4gpu.log && exit 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2n_scale_n4_wan14.json 2> gpurun_out/r2n_scale_n4.err; echo "wan14 n4 rc=$?"; tail -2 gpurun_out/r2n_scale_n4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n_scale_n4_wan14.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d.get('parity'), d['roofline']['achieved'], d.get('attn_kernel_ms_per_rank'), d.get('nvlink'))
PY
