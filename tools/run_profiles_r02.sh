# ncu evidence for profiles/ (round 2): run on the GPU box (gpurun -- bash tools/run_profiles_r02.sh).
# Reports are converted to CSV on the box and deleted (gpurun_out/ is capped at 64 MiB).  Every ncu command runs only
# after the same command has exited 0 without ncu.
O=gpurun_out
conv() { ncu -i $1 --page raw --csv > $2 2> /dev/null; }
B="python bench.py --no-cpu-baseline --sustain-s 0 --no-e2e --no-kernel-pass --no-riders --steps 1 --warmup 1"
timeout 300 python tools/bw_bench.py > $O/bw_bench_r02.txt 2>&1; echo "bw_bench rc=$?"
timeout 300 $B --config 3 > $O/plain_train_r02.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_train_r02.csv $B --config 3 > $O/ncu_lt.log 2>&1
echo "launch list train rc=$?"
timeout 300 $B --config 2 > $O/plain_infer_r02.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/launches_infer_r02.csv $B --config 2 > $O/ncu_li.log 2>&1
echo "launch list infer rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc" -s 18 -c 18 -o $O/prof_conv_infer $B --config 2 > $O/ncu_conv_infer.log 2>&1
echo "conv infer rc=$?"; conv $O/prof_conv_infer.ncu-rep $O/prof_conv_infer_raw.csv; rm -f $O/prof_conv_infer.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc" -s 35 -c 35 -o $O/prof_conv $B --config 3 > $O/ncu_conv.log 2>&1
echo "conv train rc=$?"; conv $O/prof_conv.ncu-rep $O/prof_conv_raw.csv; rm -f $O/prof_conv.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad3x3_tc" -s 18 -c 18 -o $O/prof_wgrad $B --config 3 > $O/ncu_wgrad.log 2>&1
echo "wgrad rc=$?"; conv $O/prof_wgrad.ncu-rep $O/prof_wgrad_raw.csv; rm -f $O/prof_wgrad.ncu-rep
du -sh $O; ls $O | grep -E "r02|prof_" | head -30
