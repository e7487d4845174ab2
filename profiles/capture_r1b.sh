# round-1 final-kernel evidence: GPU tests, bench lines, ncu launch list (time + DRAM bytes) and --set full capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1b.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r1b.log
timeout 600 python bench.py --workload hunyuan --steps 2 --warmup 3 > gpurun_out/bench_r1b_hunyuan.json 2> gpurun_out/bench_r1b_hunyuan.err; echo "hunyuan rc=$?"
CMD="python bench.py --workload wan13 --steps 1 --warmup 3 --no-aux --no-cpu-baseline --profile"
$CMD > gpurun_out/plain_r1b.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_launches_r1b.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vb_attn_fwd -c 3 -o gpurun_out/prof_attn_r1b $CMD > gpurun_out/ncu_full_r1b.log 2>&1
echo "full capture rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1b_n1.json 2> gpurun_out/bench_r1b_n1.err; echo "wan14 rc=$?"
ls -la gpurun_out/
