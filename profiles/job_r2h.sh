# round-2 (h): 2 GPUs — placement with uneven head counts and query-half units through the peer exchange
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "query_half or placement_table or out_heads" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "2-" 2>&1 | tail -60 > gpurun_out/r2h_pytest_2gpu.log; tail -40 gpurun_out/r2h_pytest_2gpu.log
grep -q "failed" gpurun_out/r2h_pytest_2gpu.log && exit 1
VB_ULYSSES_SPLIT=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2h_scale_n2_split.json 2> gpurun_out/r2h_scale_n2.err; echo "wan14 n2 rc=$?"; tail -3 gpurun_out/r2h_scale_n2.err
VB_ULYSSES_SPLIT=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --workload hunyuan > gpurun_out/r2h_scale_n2_hunyuan_split.json 2> gpurun_out/r2h_scale_n2_hunyuan.err; echo "hy n2 rc=$?"; tail -3 gpurun_out/r2h_scale_n2_hunyuan.err
python - <<'PY'
import json
for f in ('gpurun_out/r2h_scale_n2_split.json','gpurun_out/r2h_scale_n2_hunyuan_split.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['value'], d['e2e']['value'], d.get('parity'), d['roofline']['achieved'])
    except Exception as e:
        print(f, 'ERR', e)
PY
