set -x
timeout 200 python -m pytest tests/test_bench_contract.py tests/test_gpu_dit.py -x -q -m gpu 2>&1 | tail -4
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
