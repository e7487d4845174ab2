# round-2 (s): sliding branch lays its query rows out window group by window group (606 instead of 700 query slots at
# Wan-14B); A/B against one group per tile on the same box, full GPU suite, smoke, N=1 lines
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2s_pytest.log; tail -8 gpurun_out/r2s_pytest.log
grep -q "failed\|error" gpurun_out/r2s_pytest.log && exit 1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2s_smoke.log
VB_TAG=groups VB_SLIDING=1 timeout 200 python tests/perf_attn.py > gpurun_out/r2s_perf_groups.log 2>&1; cat gpurun_out/r2s_perf_groups.log
VB_TAG=per-tile VB_SLIDING=1 VB_ATTN_NO_WINDOW_GROUPS=1 timeout 200 python tests/perf_attn.py > gpurun_out/r2s_perf_per_tile.log 2>&1; cat gpurun_out/r2s_perf_per_tile.log
VB_TAG=groups VB_SLIDING=1 timeout 200 python tests/perf_attn.py > gpurun_out/r2s_perf_groups2.log 2>&1; cat gpurun_out/r2s_perf_groups2.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err; echo "bench rc=$?"
VB_ATTN_NO_WINDOW_GROUPS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench_n1_per_tile.json 2> gpurun_out/r2s_bench_n1_per_tile.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('r2s_bench_n1', 'r2s_bench_n1_per_tile'):
    d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
    r=d['roofline']; print(f, d['value'], d['e2e']['value'], r['achieved'], r['frac'], r['cross_attention']['ms_per_launch'], d['attn_kernel_ms_per_step'], d.get('like_for_like'))
PY
