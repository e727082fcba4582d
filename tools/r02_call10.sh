#!/bin/bash
# round 2, call 10: rows-first bilinear backward (csrc/bilinear_vh.cuh) + whole-wave grids (head_bwd, BN apply kernels):
# GPU tests, per-kernel bandwidth A/B, training-step A/B (new / old interleaved, same box)
# (historical record of the call: csrc/bilinear_vh.cuh and the MAU_BILINEAR_BWD switch existed for one commit of this round -- the rows-first kernel lost and was removed; MAU_WHOLE_WAVES=0 still works)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/r02c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02c10_pytest.log
timeout 200 python tools/bw_bench.py > $O/r02c10_bw_new.txt 2>&1; echo "bw new rc=$?"
BW_ONLY=bilinear_bwd,head_bwd,bn_bwd_apply,bn_apply_relu MAU_BILINEAR_BWD=stream MAU_WHOLE_WAVES=0 timeout 200 python tools/bw_bench.py > $O/r02c10_bw_old.txt 2>&1; echo "bw old rc=$?"
grep -E "bilinear_bwd|head_bwd|bn_bwd_apply|bn_apply_relu" $O/r02c10_bw_new.txt; echo "-- old"; cat $O/r02c10_bw_old.txt
B="python bench.py --config 3 --no-cpu-baseline --sustain-s 1 --no-e2e --no-riders"
timeout 300 $B > $O/r02c10_c3_new_a.json 2> $O/r02c10_c3_new_a.err; echo "new a rc=$?"
MAU_BILINEAR_BWD=stream MAU_WHOLE_WAVES=0 timeout 300 $B > $O/r02c10_c3_old_a.json 2> /dev/null; echo "old a rc=$?"
timeout 300 $B > $O/r02c10_c3_new_b.json 2> /dev/null; echo "new b rc=$?"
MAU_BILINEAR_BWD=stream MAU_WHOLE_WAVES=0 timeout 300 $B > $O/r02c10_c3_old_b.json 2> /dev/null; echo "old b rc=$?"
MAU_WHOLE_WAVES=0 timeout 300 $B > $O/r02c10_c3_vh_only.json 2> /dev/null; echo "vh only rc=$?"
python - <<'PY'
import json
for f in ('new_a', 'old_a', 'new_b', 'old_b', 'vh_only'):
    try:
        d = json.load(open(f'gpurun_out/r02c10_c3_{f}.json'))
        print(f, round(d['value'], 1), round(d['ms_per_step'], 3), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'], 3))
    except Exception as e:
        print(f, 'ERR', e)
PY
