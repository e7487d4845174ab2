# round-2 captures (1 GPU): --set full of one routed self-attention launch and one cross-attention launch of a Wan-14B
# step, and the launch list of a Wan-1.3B step.  The same command lines run WITHOUT ncu first (B200_PROFILING.md).
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload wan14 --steps 1 --warmup 3 --no-aux --no-cpu-baseline --profile"
CMD13="python bench.py --workload wan13 --steps 1 --warmup 3 --no-aux --no-cpu-baseline --profile"
timeout 400 $CMD > gpurun_out/r2_plain_wan14.json 2> gpurun_out/r2_plain_wan14.err && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vb_attn_fwd -c 2 -o gpurun_out/r2_prof_attn_wan14 $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "full capture rc=$?"
timeout 300 $CMD13 > gpurun_out/r2_plain_wan13.json 2> gpurun_out/r2_plain_wan13.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_wan13.csv $CMD13 > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out | grep r2_
