# round-1 (n): decoupled attention kernel — tests, bench lines, sweep, launch list + --set full capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1n.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_r1n.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1n_n1.json 2> gpurun_out/bench_r1n_n1.err; echo "wan14 rc=$?"
cut -c1-200 gpurun_out/bench_r1n_n1.json
timeout 600 python bench.py --workload hunyuan --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1n_hunyuan.json 2> gpurun_out/bench_r1n_hunyuan.err; echo "hunyuan rc=$?"
timeout 600 python tests/sweep_attn.py > gpurun_out/sweep_attn_r1n.csv 2>&1; echo "sweep rc=$?"
CMD="python bench.py --workload wan14 --steps 1 --warmup 3 --no-aux --no-cpu-baseline --profile"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vb_attn_fwd -c 2 -o gpurun_out/prof_attn_r1n_wan14 $CMD > gpurun_out/ncu_full_r1n.log 2>&1
echo "full capture rc=$?"
CMD13="python bench.py --workload wan13 --steps 1 --warmup 3 --no-aux --no-cpu-baseline --profile"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1n.csv $CMD13 > gpurun_out/ncu_launches_r1n.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/ | tail -8
