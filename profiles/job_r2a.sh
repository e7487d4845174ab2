# round-2 (a): validate the new parity tests and the reworked bench (reference arm = the reference's own processor)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/r2a_pytest.log; tail -12 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench_n1.err
cut -c1-400 gpurun_out/r2a_bench_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/r2a_ref.json
timeout 200 python tests/perf_attn.py > gpurun_out/r2a_perf.log 2>&1; cat gpurun_out/r2a_perf.log
