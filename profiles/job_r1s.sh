set -x
mkdir -p gpurun_out
L=$PWD/vorta_b200/lib/exp
: > gpurun_out/perf_attn_r1s.log
for rep in 1 2; do
for v in product p2 p3 p0; do
  if [ $v = product ]; then unset VB_LIB_PATH; else export VB_LIB_PATH=$L/libvb_$v.so; fi
  VB_QUICK=1 VB_TAG=$v timeout 60 python tests/perf_attn.py >> gpurun_out/perf_attn_r1s.log 2>&1
done
done
unset VB_LIB_PATH
cat gpurun_out/perf_attn_r1s.log
