# round-1 (f): tests, HBM-kernel table, Wan-14B bench line, launch list + --set full capture of the Wan-14B step's attention
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1f.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_r1f.log
timeout 600 python tests/perf_block.py > gpurun_out/perf_block_r1f.csv 2> gpurun_out/perf_block_r1f.err; echo "perf_block rc=$?"
cat gpurun_out/perf_block_r1f.csv; tail -5 gpurun_out/perf_block_r1f.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1f_n1.json 2> gpurun_out/bench_r1f_n1.err; echo "wan14 rc=$?"
cut -c1-600 gpurun_out/bench_r1f_n1.json
CMD="python bench.py --workload wan14 --steps 1 --warmup 3 --no-aux --no-cpu-baseline --profile"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vb_attn_fwd -c 4 -o gpurun_out/prof_attn_r1f_wan14 $CMD > gpurun_out/ncu_full_r1f.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/ | tail -8
