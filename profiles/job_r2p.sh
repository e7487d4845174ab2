# round-2 (p): K / V exchange stores overlapped with the next projection GEMM (side stream, capped grid) — 2 GPUs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "2-" 2>&1 | tail -30 > gpurun_out/r2p_pytest_2gpu.log; tail -6 gpurun_out/r2p_pytest_2gpu.log
grep -q "failed" gpurun_out/r2p_pytest_2gpu.log && exit 1
for ov in 1 0 1 0; do
VB_ULYSSES_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2p_scale_n2_overlap$ov.json 2> gpurun_out/r2p_scale_n2_overlap$ov.err; echo "overlap=$ov rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2p_scale_n2_overlap$ov.json') if l.startswith('{')][-1])
print('overlap=$ov', d['value'], d['e2e']['value'], d['parity']['equal'], d['attn_kernel_ms_per_rank'])
PY
done
