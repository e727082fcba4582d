#!/bin/bash
# last 1-GPU call of round 2: lean bilinear backward (csrc/bilinear_bwd_lean.cuh) against the first-generation kernel
# (MAU_BILINEAR_BWD=stream), kernel level and step level, interleaved on one box; then -- with whichever won -- smoke, the
# default bench line, the bandwidth table and the launch list of one training step.
# (historical record of the call: MAU_BILINEAR_BWD=stream selected the first-generation kernel at that commit; it was removed
# after this call, the switch no longer exists)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02g_pytest.log
BW_ONLY=bilinear_bwd timeout 100 python tools/bw_bench.py > $O/r02g_bw_lean.txt 2>&1
BW_ONLY=bilinear_bwd MAU_BILINEAR_BWD=stream timeout 100 python tools/bw_bench.py > $O/r02g_bw_first.txt 2>&1
echo "-- lean"; cat $O/r02g_bw_lean.txt; echo "-- first generation"; cat $O/r02g_bw_first.txt
B="python bench.py --config 3 --no-cpu-baseline --sustain-s 1 --no-e2e --no-riders"
timeout 300 $B > $O/r02g_c3_lean_a.json 2> $O/r02g_c3_lean_a.err; echo "lean a rc=$?"
MAU_BILINEAR_BWD=stream timeout 300 $B > $O/r02g_c3_first_a.json 2> /dev/null; echo "first a rc=$?"
timeout 300 $B > $O/r02g_c3_lean_b.json 2> /dev/null; echo "lean b rc=$?"
MAU_BILINEAR_BWD=stream timeout 300 $B > $O/r02g_c3_first_b.json 2> /dev/null; echo "first b rc=$?"
WIN=$(python - <<'PY'
import json
def ms(tag):
    v = []
    for s in 'ab':
        try:
            d = json.load(open(f'gpurun_out/r02g_c3_{tag}_{s}.json'))
            v.append((d['ms_per_step'], d['sustained']['ms_per_step']))
        except Exception:
            pass
    return v
L, F = ms('lean'), ms('first')
import sys
print('lean', L, 'first', F, file=sys.stderr)
score = lambda v: sum(a + b for a, b in v) / max(len(v), 1) if v else 1e9
print('lean' if score(L) <= score(F) else 'stream')
PY
)
echo "winner: $WIN"
if [ "$WIN" = "stream" ]; then
  export MAU_BILINEAR_BWD=stream
  timeout 600 python -m pytest tests -m gpu -q -k "bilinear or bf16 or train or smoke or full_size" > $O/r02g_pytest_first.log 2>&1; echo "pytest (first-generation kernel) rc=$?"; tail -2 $O/r02g_pytest_first.log
fi
python __graft_entry__.py smoke > $O/r02g_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02g_smoke.log
timeout 600 python bench.py > $O/r02g_bench_default.json 2> $O/r02g_bench_default.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02g_bench_default.json'))
print(d['metric'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d['e2e'], 'sust', d['sustained']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['burst']['frac'], 'cpu', d['cpu_baseline']['value'])
print('inference', round(d['inference']['value'],1), d['inference']['e2e'])
for k,v in d['riders'].items(): print(k, round(v['value'],1))
"
timeout 300 python tools/bw_bench.py > $O/bw_bench_r02.txt 2>&1; echo "bw_bench rc=$?"
P="python bench.py --no-cpu-baseline --sustain-s 0 --no-e2e --no-kernel-pass --no-riders --steps 1 --warmup 1"
timeout 300 $P --config 3 > $O/plain_train_r02.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_train_r02.csv $P --config 3 > $O/ncu_lt.log 2>&1
echo "launch list train rc=$?"
