# round-2 (d): HunyuanVideo fused prologue, dense SP baseline, hygiene changes; HY step time before / after
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2d_pytest.log; tail -6 gpurun_out/r2d_pytest.log
grep -q "failed" gpurun_out/r2d_pytest.log && exit 1
timeout 600 python bench.py --workload hunyuan --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_n1_hunyuan.json 2> gpurun_out/r2d_bench_n1_hunyuan.err; echo "hy bench rc=$?"; tail -3 gpurun_out/r2d_bench_n1_hunyuan.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench_n1_hunyuan.json'))
print(d['value'], d['e2e']['value'], d['attn_kernel_ms_per_step'], d['attn_flops_per_step'], d['routing_mix'])
r=d['roofline']; print(r['achieved'], r['ms_per_launch'])
PY
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench_n1.json'))
print(d['value'], d['attn_flops_per_step'], d['routing_mix'])
PY
