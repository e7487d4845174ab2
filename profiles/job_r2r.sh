# round-2 (r): 8 GPUs, same box — query parts (halves only vs halves + quarters) and the branch cost model (Wan-14B)
set -x
mkdir -p gpurun_out
i=0
for cfg in "2|1.0,1.12,1.25" "2,4|1.0,1.12,1.25" "2|1.0,1.06,1.33" "2,4|1.0,1.06,1.33"; do
i=$((i+1))
ways=${cfg%%|*}; costs=${cfg##*|}
VB_ULYSSES_SPLIT_WAYS=$ways VB_ULYSSES_COSTS=$costs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2r_scale_n8_run$i.json 2> gpurun_out/r2r_scale_n8_run$i.err; echo "ways=$ways costs=$costs rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2r_scale_n8_run$i.json') if l.startswith('{')][-1])
print('ways=$ways costs=$costs', d['value'], d['e2e']['value'], d['parity']['equal'], max(d['attn_kernel_ms_per_rank']), d['attn_kernel_ms_per_rank'])
PY
done
