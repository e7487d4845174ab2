set -x
mkdir -p gpurun_out
L=$PWD/vorta_b200/lib/exp
VB_LIB_PATH=$L/libvb_dec.so timeout 90 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense_attention or wan_branches" 2>&1 | tail -3
VB_LIB_PATH=$L/libvb_dec.so VB_QUICK=1 VB_TAG=dec timeout 60 python tests/perf_attn.py 2>&1 | tail -2
VB_LIB_PATH=$L/libvb_dece.so VB_QUICK=1 VB_TAG=dec-early timeout 60 python tests/perf_attn.py 2>&1 | tail -2
VB_QUICK=1 VB_TAG=product timeout 60 python tests/perf_attn.py 2>&1 | tail -2
VB_LIB_PATH=$L/libvb_dectl.so timeout 60 python tests/timeline_dec.py > gpurun_out/timeline_dec_r1l.log 2>&1; echo "tl rc=$?"
sed -n 1,30p gpurun_out/timeline_dec_r1l.log; tail -4 gpurun_out/timeline_dec_r1l.log
