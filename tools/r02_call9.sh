# (historical: at that commit the second stream forked BEHIND the data gradient by default and MAU_WGRAD_FORK_EARLY=1 selected the early fork, which won and is the default now; MAU_WGRAD_FORK_LATE=1 selects the late fork)
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/r02c9_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02c9_pytest.log
B="python bench.py --config 3 --no-cpu-baseline --sustain-s 1 --no-e2e"
timeout 300 $B > $O/r02c9_bench_c3.json 2> $O/r02c9_bench_c3.err; echo "bench rc=$?"
MAU_WGRAD_FORK_EARLY=1 timeout 300 $B > $O/r02c9_bench_c3_forkearly.json 2> /dev/null; echo "fork-early rc=$?"
MAU_FLAGS=16384 timeout 300 $B > $O/r02c9_bench_c3_nostats.json 2> /dev/null; echo "no-conv-stats rc=$?"
timeout 300 $B > $O/r02c9_bench_c3_b.json 2> /dev/null; echo "bench again rc=$?"
timeout 300 python bench.py --config 4 --no-cpu-baseline --sustain-s 0 --no-e2e > $O/r02c9_bench_c4.json 2> /dev/null; echo "c4 rc=$?"
timeout 300 python bench.py --config 2 --no-cpu-baseline --sustain-s 0 > $O/r02c9_bench_c2.json 2> /dev/null; echo "c2 rc=$?"
timeout 300 python tools/timeline.py --config 3 --tag r02c9 > $O/r02c9_timeline.txt 2>&1; grep -A14 "step span" $O/r02c9_timeline.txt
python -c "
import json
for f in ('r02c9_bench_c3','r02c9_bench_c3_forkearly','r02c9_bench_c3_nostats','r02c9_bench_c3_b','r02c9_bench_c4','r02c9_bench_c2'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'],3), 'e2e', d.get('e2e'))
    except Exception as e: print(f, 'ERR', e)
"
