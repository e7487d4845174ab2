set -x
mkdir -p gpurun_out
L=$PWD/vorta_b200/lib/exp
: > gpurun_out/perf_attn_r1m.log
for v in dece dece_p0 dece_p1 dece_p3; do
  VB_LIB_PATH=$L/libvb_$v.so VB_QUICK=1 VB_TAG=$v timeout 60 python tests/perf_attn.py >> gpurun_out/perf_attn_r1m.log 2>&1; echo "$v rc=$?"
done
VB_LIB_PATH=$L/libvb_dece.so VB_TAG=dece timeout 120 python tests/perf_attn.py >> gpurun_out/perf_attn_r1m.log 2>&1
VB_TAG=product timeout 120 python tests/perf_attn.py >> gpurun_out/perf_attn_r1m.log 2>&1
cat gpurun_out/perf_attn_r1m.log
VB_LIB_PATH=$L/libvb_dece.so timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
