# round-2 (t): 2 GPUs with the window-group sliding schedule — five exchange modes bit-identical to 1 GPU, bench lines
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "2-" 2>&1 | tail -60 > gpurun_out/r2t_pytest_2gpu.log; tail -12 gpurun_out/r2t_pytest_2gpu.log
grep -q "failed" gpurun_out/r2t_pytest_2gpu.log && exit 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2t_scale_n2_wan14.json 2> gpurun_out/r2t_scale_n2.err; echo "wan14 n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload hunyuan > gpurun_out/r2t_scale_n2_hunyuan.json 2> gpurun_out/r2t_scale_n2_hunyuan.err; echo "hy n2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2t_scale_n2_wan14.json','gpurun_out/r2t_scale_n2_hunyuan.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['value'], d['e2e']['value'], d.get('parity'), d['roofline']['achieved'], d.get('attn_kernel_ms_per_rank'))
    except Exception as e:
        print(f, 'ERR', e)
PY
