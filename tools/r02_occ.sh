#!/bin/bash
# round 2, last experiment: lean bilinear backward at three resident blocks per SM (80 registers) against the
# 127-register build (MAU_BILINEAR_OCC=2): tests on the default, kernel level and step level A/B, smoke
# (historical record of the call: the 127-register build and MAU_BILINEAR_OCC existed at that commit only)
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/r02o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02o_pytest.log
BW_ONLY=bilinear_bwd timeout 100 python tools/bw_bench.py > $O/r02o_bw_occ3.txt 2>&1
BW_ONLY=bilinear_bwd MAU_BILINEAR_OCC=2 timeout 100 python tools/bw_bench.py > $O/r02o_bw_occ2.txt 2>&1
echo "-- 3 blocks/SM"; cat $O/r02o_bw_occ3.txt; echo "-- 2 blocks/SM"; cat $O/r02o_bw_occ2.txt
B="python bench.py --config 3 --no-cpu-baseline --sustain-s 1 --no-e2e --no-riders"
timeout 300 $B > $O/r02o_c3_occ3.json 2> /dev/null; echo "occ3 rc=$?"
MAU_BILINEAR_OCC=2 timeout 300 $B > $O/r02o_c3_occ2.json 2> /dev/null; echo "occ2 rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02o_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02o_smoke.log
python -c "
import json
for f in ('occ3','occ2'):
    d=json.load(open('gpurun_out/r02o_c3_%s.json' % f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), d['sustained']['ms_per_step'])"
