# round-2: compute-sanitizer racecheck (ONE tool per call, B200_PROFILING.md) on the smallest case that runs every kernel
# of the path: __graft_entry__.smoke() (router, selection, gathers, the tcgen05 attention kernel in all three branches)
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 50 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_racecheck_smoke.log 2>&1
echo "racecheck rc=$?"
tail -30 gpurun_out/r2_racecheck_smoke.log
