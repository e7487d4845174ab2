set -x
CMD="python bench.py --workload wan13 --steps 1 --warmup 3 --no-aux --no-cpu-baseline --profile"
$CMD > gpurun_out/plain_r1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches_r1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_r1.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vb_attn_fwd -c 3 -o gpurun_out/prof_attn_r1 $CMD > gpurun_out/ncu_full_r1.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
