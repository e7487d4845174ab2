set -x
mkdir -p gpurun_out
L=$PWD/vorta_b200/lib/exp
export VB_LIB_PATH=$L/libvb_dec.so
timeout 90 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense_attention or wan_branches" 2>&1 | tail -5
echo "parity-small rc=$?"
VB_QUICK=1 VB_TAG=dec timeout 60 python tests/perf_attn.py > gpurun_out/perf_attn_r1i.log 2>&1; echo "perf rc=$?"
cat gpurun_out/perf_attn_r1i.log
unset VB_LIB_PATH
VB_QUICK=1 VB_TAG=product timeout 60 python tests/perf_attn.py >> gpurun_out/perf_attn_r1i.log 2>&1; tail -2 gpurun_out/perf_attn_r1i.log
