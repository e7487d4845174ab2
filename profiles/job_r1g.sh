set -x
mkdir -p gpurun_out
L=$PWD/vorta_b200/lib/exp
VB_LIB_PATH=$L/libvb_timeline.so timeout 300 python tests/timeline_attn.py > gpurun_out/timeline_r1g.log 2>&1; echo "timeline rc=$?"
cat gpurun_out/timeline_r1g.log | tail -45
VB_QUICK=1 VB_TAG=product timeout 300 python tests/perf_attn.py > gpurun_out/perf_attn_r1g.log 2>&1
VB_QUICK=1 VB_TAG=nosoftmax VB_LIB_PATH=$L/libvb_nosoftmax.so timeout 300 python tests/perf_attn.py >> gpurun_out/perf_attn_r1g.log 2>&1
VB_QUICK=1 VB_TAG=nomufu VB_LIB_PATH=$L/libvb_nomufu.so timeout 300 python tests/perf_attn.py >> gpurun_out/perf_attn_r1g.log 2>&1
cat gpurun_out/perf_attn_r1g.log
timeout 600 python tests/perf_block.py > gpurun_out/perf_block_r1g.csv 2> gpurun_out/perf_block_r1g.err; grep select gpurun_out/perf_block_r1g.csv
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "selection or coreset or pool" 2>&1 | tail -3
