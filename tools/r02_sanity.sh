#!/bin/bash
# last sanity pass of round 2: GPU tests + smoke at HEAD, and the bandwidth table of the final kernels
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02s_smoke.log
timeout 200 python bench.py --config 3 --no-cpu-baseline --sustain-s 1 --no-e2e --no-riders > gpurun_out/r02s_c3.json 2> /dev/null; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02s_c3.json')); print(round(d['value'],1), round(d['ms_per_step'],3), d['sustained']['ms_per_step'])"
# ncu --set full of the bandwidth kernels this call's changes touched (after the same command exited 0 without ncu)
export BW_ONLY=bilinear_bwd,head_bwd,bn_bwd_apply BW_ITERS=1 BW_SETS=1
timeout 200 python tools/bw_bench.py > gpurun_out/bw_plain3.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bilinear_bwd|head_bwd|bn_bwd_apply" -c 14 -o gpurun_out/prof_bw3 python tools/bw_bench.py > gpurun_out/ncu_bw3.log 2>&1
echo "ncu bw rc=$?"
ncu -i gpurun_out/prof_bw3.ncu-rep --page raw --csv > gpurun_out/prof_bw3_raw.csv 2> /dev/null; rm -f gpurun_out/prof_bw3.ncu-rep
unset BW_ONLY BW_ITERS BW_SETS
ls -la gpurun_out/prof_bw3_raw.csv
