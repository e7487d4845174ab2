set -x
mkdir -p gpurun_out
L=$PWD/vorta_b200/lib/exp
: > gpurun_out/perf_attn_r1h.log
for v in product nodbg spinmma spinsm spinboth smr smrspin; do
  if [ $v = product ]; then unset VB_LIB_PATH; else export VB_LIB_PATH=$L/libvb_$v.so; fi
  VB_QUICK=1 VB_TAG=$v timeout 120 python tests/perf_attn.py >> gpurun_out/perf_attn_r1h.log 2>&1; echo "$v rc=$?"
done
cat gpurun_out/perf_attn_r1h.log
VB_LIB_PATH=$L/libvb_smrspin.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "branches_vs_oracle or wan_branches or dense_attention" 2>&1 | tail -3
