set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1c.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_r1c.log
timeout 600 python tests/perf_block.py > gpurun_out/perf_block_r1c.csv 2> gpurun_out/perf_block_r1c.err; echo "perf_block rc=$?"
cat gpurun_out/perf_block_r1c.csv; tail -5 gpurun_out/perf_block_r1c.err
timeout 600 python bench.py --workload wan13 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1c_wan13.json 2> gpurun_out/bench_r1c_wan13.err; echo "wan13 rc=$?"
VB_ATTN_SPLIT_LAUNCHES=1 timeout 600 python bench.py --workload wan13 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1c_wan13_split.json 2> gpurun_out/bench_r1c_wan13_split.err; echo "wan13 split rc=$?"
cat gpurun_out/bench_r1c_wan13.json gpurun_out/bench_r1c_wan13_split.json | cut -c1-1500
