# round-2 (q): final kernels;: unit placement (uneven head counts + query halves);: 8 GPUs — Wan-14B and HunyuanVideo (BASELINE configs[2], [3]) with the in-bench bit-parity check
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2q_scale_n8.json 2> gpurun_out/r2q_scale_n8.err; echo "wan14 n8 rc=$?"; tail -3 gpurun_out/r2q_scale_n8.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 --workload hunyuan > gpurun_out/r2q_scale_n8_hunyuan.json 2> gpurun_out/r2q_scale_n8_hunyuan.err; echo "hy n8 rc=$?"; tail -3 gpurun_out/r2q_scale_n8_hunyuan.err
python - <<'PY'
import json
for f in ('gpurun_out/r2q_scale_n8.json','gpurun_out/r2q_scale_n8_hunyuan.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['value'], d['e2e']['value'], d.get('parity'), d['roofline']['achieved'], d.get('attn_kernel_ms_per_rank'), d.get('nvlink'))
    except Exception as e:
        print(f, 'ERR', e)
PY
