O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wgrad" > $O/r02c3_pytest_wgrad.log 2>&1; echo "wgrad tests rc=$?"
tail -12 $O/r02c3_pytest_wgrad.log
MAU_WGRAD_PAIR=0 timeout 900 python -m pytest tests -m gpu -q > $O/r02c3_pytest_nopair.log 2>&1; echo "pytest (pair off) rc=$?"
tail -6 $O/r02c3_pytest_nopair.log
timeout 900 python -m pytest tests -m gpu -q -k "not wgrad_against" > $O/r02c3_pytest_pair.log 2>&1; echo "pytest (pair on) rc=$?"
tail -6 $O/r02c3_pytest_pair.log
timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 0 --profile-layers > $O/r02c3_bench_c3_pair.json 2> $O/r02c3_layers_c3_pair.txt; echo "bench pair rc=$?"
MAU_WGRAD_PAIR=0 timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 0 > $O/r02c3_bench_c3_nopair.json 2> /dev/null; echo "bench nopair rc=$?"
python -c "
import json
for f in ('r02c3_bench_c3_pair','r02c3_bench_c3_nopair'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3))
    except Exception as e: print(f, 'ERR', e)
"
grep -E "wgrad" $O/r02c3_layers_c3_pair.txt | head -20
