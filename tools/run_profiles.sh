# ncu evidence for profiles/: run on the GPU box (gpurun -- bash tools/run_profiles.sh).
# Reports are converted to CSV on the box and deleted (gpurun_out/ is capped at 64 MiB).
set -x
O=gpurun_out
conv() { ncu -i $1 --page raw --csv > $2 2> /dev/null; }
export BW_ONLY=bilinear,bilinear_bwd,head,head_bwd,bn_stats,bn_apply_relu,bn_bwd_reduce,bn_bwd_apply,maxpool,maxpool_bwd,nchw_to_nhwc,embed_broadcast BW_ITERS=1 BW_SETS=1
timeout 300 python tools/bw_bench.py > $O/bw_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bilinear|head_|bn_|maxpool|nchw_to_nhwc|embed_broadcast" -c 46 -o $O/prof_bw python tools/bw_bench.py > $O/ncu_bw.log 2>&1
echo bw rc=$?
conv $O/prof_bw.ncu-rep $O/prof_bw_raw.csv
rm -f $O/prof_bw.ncu-rep
unset BW_ONLY BW_ITERS BW_SETS
timeout 300 python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > $O/plain_train2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches_train_r01b.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_lt.log 2>&1
echo lt rc=$?
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/plain_infer2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_infer_r01b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_li.log 2>&1
echo li rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc_v2" -s 18 -c 18 -o $O/prof_conv_infer python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_conv_infer.log 2>&1
echo conv_infer rc=$?
conv $O/prof_conv_infer.ncu-rep $O/prof_conv_infer_raw.csv
rm -f $O/prof_conv_infer.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc_v2" -s 53 -c 35 -o $O/prof_conv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_conv.log 2>&1
echo conv rc=$?
conv $O/prof_conv.ncu-rep $O/prof_conv_raw.csv
rm -f $O/prof_conv.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad3x3_tc_v2" -s 18 -c 18 -o $O/prof_wgrad python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_wgrad.log 2>&1
echo wgrad rc=$?
conv $O/prof_wgrad.ncu-rep $O/prof_wgrad_raw.csv
rm -f $O/prof_wgrad.ncu-rep
du -sh $O; ls -la $O
