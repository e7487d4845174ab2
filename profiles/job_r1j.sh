set -x
mkdir -p gpurun_out
L=$PWD/vorta_b200/lib/exp
VB_LIB_PATH=$L/libvb_dectl.so timeout 60 python tests/timeline_dec.py > gpurun_out/timeline_dec_r1j.log 2>&1; echo "tl rc=$?"
cat gpurun_out/timeline_dec_r1j.log
VB_LIB_PATH=$L/libvb_dece.so VB_QUICK=1 VB_TAG=dec-early timeout 60 python tests/perf_attn.py 2>&1 | tail -2
