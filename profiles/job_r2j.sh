# round-2 (j): coreset selection with 8 lanes per row: 3 vs 2 CTAs per SM; HBM-bound kernels incl. the HunyuanVideo prologue
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "coreset or selection or fullsize or pool_unpool or full_size_properties" 2>&1 | tail -5
timeout 300 python tests/perf_block.py > gpurun_out/r2j_hbm_kernels_select3.csv 2>&1; cat gpurun_out/r2j_hbm_kernels_select3.csv
VB_LIB_PATH=$PWD/vorta_b200/lib/exp/libvb_select2.so timeout 300 python tests/perf_block.py > gpurun_out/r2j_hbm_kernels_select2.csv 2>&1; grep select gpurun_out/r2j_hbm_kernels_select2.csv
