O=gpurun_out
timeout 120 python -m pytest tests/test_bf16_layers_gpu.py -m gpu -q -x > $O/r02c6_pytest_layers.log 2>&1; echo "layers rc=$?"
tail -4 $O/r02c6_pytest_layers.log
timeout 900 python -m pytest tests -m gpu -q > $O/r02c6_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/r02c6_pytest.log
timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 1 > $O/r02c6_bench_c3.json 2> $O/r02c6_bench_c3.err; echo "bench rc=$?"
MAU_FLAGS=16384 timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 1 > $O/r02c6_bench_c3_nostats.json 2> /dev/null; echo "bench no-conv-stats rc=$?"
timeout 300 python bench.py --config 4 --no-cpu-baseline --sustain-s 0 > $O/r02c6_bench_c4.json 2> /dev/null; echo "bench c4 rc=$?"
timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 0 --profile-layers > /dev/null 2> $O/r02c6_layers_c3.txt; echo "layers rc=$?"
python -c "
import json
for f in ('r02c6_bench_c3','r02c6_bench_c3_nostats','r02c6_bench_c4'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
"
