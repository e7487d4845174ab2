# round-2 (v): final single-GPU validation with the window-group schedule — all GPU tests, smoke, the default bench line
# with its CPU legs, the reference arm, and a fresh ncu launch list of a Wan-1.3B step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -25 > gpurun_out/r2v_pytest_gpu.log; tail -8 gpurun_out/r2v_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2v_bench_n1_wan14.json 2> gpurun_out/r2v_bench_n1_wan14.err; echo "bench rc=$?"; tail -3 gpurun_out/r2v_bench_n1_wan14.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2v_reference_arm.json 2> gpurun_out/r2v_reference_arm.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2v_bench_n1_wan14.json'))
print(d['value'], d['e2e'], d['attn_flops_per_step']['library_counter']==d['attn_flops_per_step']['closed_form'])
r=d['roofline']; print(r['achieved'], r['frac'], r['ms_per_launch'], r['cross_attention']['achieved'], r['cross_attention']['ms_per_launch'], r['cross_attention']['hbm_gbs'])
print(d['like_for_like']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores']); print(d['aux']); print(d['clocks'])
r=json.load(open('gpurun_out/r2v_reference_arm.json')); print(r['value'], r['cpu_baseline']['cores'])
PY
CMD13="python bench.py --workload wan13 --steps 1 --warmup 3 --no-aux --no-cpu-baseline --profile"
timeout 300 $CMD13 > gpurun_out/r2v_plain_wan13.json 2> gpurun_out/r2v_plain_wan13.err && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2v_launches_wan13.csv $CMD13 > gpurun_out/r2v_ncu_launches.log 2>&1
echo "launch list rc=$?"
