# round-2 (m): blend (Train) mode reworked — fp32 accumulation, device-resident scores, one launch per branch for all batches
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2m_pytest_gpu.log; tail -8 gpurun_out/r2m_pytest_gpu.log
grep -q "failed" gpurun_out/r2m_pytest_gpu.log && exit 1
timeout 400 python tests/perf_ab.py > gpurun_out/r2m_perf_ab.log 2>&1; cat gpurun_out/r2m_perf_ab.log
