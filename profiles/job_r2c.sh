# round-2 (c): dynamic work-item queue (greedy LPT), unrolled PV fast path, split only for all-single schedules
set -x
mkdir -p gpurun_out
W=$PWD/vorta_b200/lib/exp/libvb_watchdog.so
VB_LIB_PATH=$W timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not full_size and not fullsize" 2>&1 | tail -15 > gpurun_out/r2c_pytest_watchdog.log; tail -4 gpurun_out/r2c_pytest_watchdog.log
grep -q "passed" gpurun_out/r2c_pytest_watchdog.log || exit 1
grep -q "failed" gpurun_out/r2c_pytest_watchdog.log && exit 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest.log; tail -6 gpurun_out/r2c_pytest.log
grep -q "failed" gpurun_out/r2c_pytest.log && exit 1
timeout 400 python tests/perf_ab.py > gpurun_out/r2c_perf_ab.log 2>&1; cat gpurun_out/r2c_perf_ab.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench_n1.json'))
print(d['value'], d['e2e']['value'], d['attn_flops_per_step'])
r=d['roofline']; print(r['achieved'], r['ms_per_launch'], r['cross_attention']['achieved'], r['cross_attention']['ms_per_launch'])
print(d['aux'])
PY
