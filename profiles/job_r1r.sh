set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
VB_TAG=uniform-warp timeout 120 python tests/perf_attn.py > gpurun_out/perf_attn_r1r.log 2>&1; cat gpurun_out/perf_attn_r1r.log
