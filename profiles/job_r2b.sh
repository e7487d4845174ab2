# round-2 (b): first run of the persistent attention kernel (work loop, narrow tail MMAs, split items)
set -x
mkdir -p gpurun_out
W=$PWD/vorta_b200/lib/exp/libvb_watchdog.so
# 1) protocol bugs become traps, not hangs: watchdog build, small cases first
VB_LIB_PATH=$W timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not full_size and not fullsize" 2>&1 | tail -15 > gpurun_out/r2b_pytest_watchdog.log; tail -8 gpurun_out/r2b_pytest_watchdog.log
grep -q "passed" gpurun_out/r2b_pytest_watchdog.log || exit 1
grep -q "failed" gpurun_out/r2b_pytest_watchdog.log && exit 1
# 2) product build: all GPU tests
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest.log; tail -8 gpurun_out/r2b_pytest.log
grep -q "failed" gpurun_out/r2b_pytest.log && exit 1
# 3) isolated rates: persistent grid vs one CTA per item (the round-1 schedule), split items on / off
VB_TAG=persistent timeout 200 python tests/perf_attn.py > gpurun_out/r2b_perf_persistent.log 2>&1; cat gpurun_out/r2b_perf_persistent.log
VB_TAG=per-item VB_ATTN_GRID=items timeout 200 python tests/perf_attn.py > gpurun_out/r2b_perf_items.log 2>&1; cat gpurun_out/r2b_perf_items.log
VB_TAG=no-split VB_ATTN_NO_SPLIT=1 VB_QUICK2=1 timeout 200 python tests/perf_attn.py > gpurun_out/r2b_perf_nosplit.log 2>&1; cat gpurun_out/r2b_perf_nosplit.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_n1.json'))
print(d['value'], d['e2e']['value'], d['attn_flops_per_step'])
r=d['roofline']; print(r['achieved'], r['ms_per_launch'], r['cross_attention']['achieved'], r['cross_attention']['ms_per_launch'])
print(d['aux'])
PY
