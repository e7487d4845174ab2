set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r1q.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r1q.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1q_default.json 2> gpurun_out/bench_r1q_default.err; echo "bench rc=$?"
cut -c1-260 gpurun_out/bench_r1q_default.json; tail -2 gpurun_out/bench_r1q_default.err
