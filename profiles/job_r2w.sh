# round-2 (w): the frozen library (window-group lookup through a map) once more through the GPU suite and smoke
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2w_pytest_gpu.log; tail -3 gpurun_out/r2w_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2w_smoke.log
