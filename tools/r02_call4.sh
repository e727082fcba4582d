O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02c4_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $O/r02c4_pytest.log
timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 1 > $O/r02c4_bench_c3.json 2> $O/r02c4_bench_c3.err; echo "bench rc=$?"
MAU_FLAGS=8192 timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 1 > $O/r02c4_bench_c3_nooverlap.json 2> /dev/null; echo "bench no-overlap rc=$?"
timeout 300 python bench.py --config 4 --no-cpu-baseline --sustain-s 0 > $O/r02c4_bench_c4.json 2> /dev/null; echo "bench c4 rc=$?"
MAU_FLAGS=8192 timeout 300 python bench.py --config 4 --no-cpu-baseline --sustain-s 0 > $O/r02c4_bench_c4_nooverlap.json 2> /dev/null; echo "bench c4 no-overlap rc=$?"
timeout 300 python bench.py --config 3 --criterion l1-gradient-ssim --no-cpu-baseline --sustain-s 0 > $O/r02c4_bench_c3_ssim.json 2> /dev/null; echo "ssim rc=$?"
python -c "
import json
for f in ('r02c4_bench_c3','r02c4_bench_c3_nooverlap','r02c4_bench_c4','r02c4_bench_c4_nooverlap','r02c4_bench_c3_ssim'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
"
