# round-2 (e): sliding heads read the raster tensors directly (5-D TMA boxes, no tile-major copy); HY fused prologue
set -x
mkdir -p gpurun_out
W=$PWD/vorta_b200/lib/exp/libvb_watchdog.so
VB_LIB_PATH=$W timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not fullsize" 2>&1 | tail -15 > gpurun_out/r2e_pytest_watchdog.log; tail -4 gpurun_out/r2e_pytest_watchdog.log
grep -q "passed" gpurun_out/r2e_pytest_watchdog.log || exit 1
grep -q "failed" gpurun_out/r2e_pytest_watchdog.log && exit 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2e_pytest.log; tail -6 gpurun_out/r2e_pytest.log
grep -q "failed" gpurun_out/r2e_pytest.log && exit 1
./tests/micro/gather4_probe > gpurun_out/r2e_gather4_probe.log 2>&1; cat gpurun_out/r2e_gather4_probe.log
timeout 400 python tests/perf_ab.py > gpurun_out/r2e_perf_ab.log 2>&1; cat gpurun_out/r2e_perf_ab.log
VB_TAG=direct timeout 200 python tests/perf_attn.py > gpurun_out/r2e_perf_direct.log 2>&1; grep -i "sliding\|coreset" gpurun_out/r2e_perf_direct.log
VB_TAG=gather VB_ATTN_SLIDING_GATHER=1 timeout 200 python tests/perf_attn.py > gpurun_out/r2e_perf_gather.log 2>&1; grep -i "sliding" gpurun_out/r2e_perf_gather.log
timeout 600 python bench.py --workload hunyuan --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_n1_hunyuan.json 2> gpurun_out/r2e_bench_n1_hunyuan.err; echo "hy bench rc=$?"; tail -3 gpurun_out/r2e_bench_n1_hunyuan.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_n1_hunyuan.json'))
print(d['value'], d['e2e']['value'], d['attn_kernel_ms_per_step'], d['attn_flops_per_step'], d['routing_mix'])
r=d['roofline']; print(r['achieved'], r['ms_per_launch'])
PY
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_n1.json'))
print(d['value'], d['attn_flops_per_step'], d['routing_mix'], d['gpu_launches'])
r=d['roofline']; print(r['achieved'], r['ms_per_launch'], r['cross_attention']['ms_per_launch'])
print(d['aux'])
PY
