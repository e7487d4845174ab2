# round-2 (l): final 2-GPU validation — 4 exchange modes bit-identical to 1 GPU, bench lines, NVLink byte counters
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "2-" 2>&1 | tail -60 > gpurun_out/r2l_pytest_2gpu.log; tail -12 gpurun_out/r2l_pytest_2gpu.log
grep -q "failed" gpurun_out/r2l_pytest_2gpu.log && exit 1
nvidia-smi nvlink -gt d > gpurun_out/r2l_nvlink_before.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2l_scale_n2_wan14.json 2> gpurun_out/r2l_scale_n2.err; echo "wan14 n2 rc=$?"
nvidia-smi nvlink -gt d > gpurun_out/r2l_nvlink_after.txt 2>&1
# forwards in the window: 1 routing probe + 3 warm-up + 5 timed + (1 + 5) e2e = 15, plus the one-layer parity check
python profiles/nvlink_diff.py gpurun_out/r2l_nvlink_before.txt gpurun_out/r2l_nvlink_after.txt 15 | tee gpurun_out/r2l_nvlink_wan14_n2.csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload hunyuan > gpurun_out/r2l_scale_n2_hunyuan.json 2> gpurun_out/r2l_scale_n2_hunyuan.err; echo "hy n2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2l_scale_n2_wan14.json','gpurun_out/r2l_scale_n2_hunyuan.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['value'], d['e2e']['value'], d.get('parity'), d['roofline']['achieved'], d.get('attn_kernel_ms_per_rank'))
    except Exception as e:
        print(f, 'ERR', e)
PY
