# round-1 (t): FINAL kernels — bench line, sweep, --set full capture (Wan-14B step) and launch list (Wan-1.3B step)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1t_n1.json 2> gpurun_out/bench_r1t_n1.err; echo "wan14 rc=$?"
cut -c1-200 gpurun_out/bench_r1t_n1.json
timeout 300 python tests/sweep_attn.py > gpurun_out/sweep_attn_r1t.log 2>&1; echo "sweep rc=$?"; cp gpurun_out/sweep_attn.csv gpurun_out/sweep_attn_r1t.csv
CMD="python bench.py --workload wan14 --steps 1 --warmup 3 --no-aux --no-cpu-baseline --profile"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:vb_attn_fwd -c 2 -o gpurun_out/prof_attn_r1t_wan14 $CMD > gpurun_out/ncu_full_r1t.log 2>&1
echo "full capture rc=$?"
CMD13="python bench.py --workload wan13 --steps 1 --warmup 3 --no-aux --no-cpu-baseline --profile"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1t.csv $CMD13 > gpurun_out/ncu_launches_r1t.log 2>&1
echo "launch list rc=$?"
VB_LIB_PATH=$PWD/vorta_b200/lib/exp/libvb_timeline.so timeout 60 python tests/timeline_attn.py > gpurun_out/timeline_r1t.log 2>&1; tail -4 gpurun_out/timeline_r1t.log
