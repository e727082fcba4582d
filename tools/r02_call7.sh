O=gpurun_out
timeout 200 python -m pytest tests/test_bf16_layers_gpu.py tests/test_gpu_parity.py -m gpu -q -x -k "teacher or train_step or full_size_training or kat_train or unusual" > $O/r02c7_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/r02c7_pytest.log
timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 1 > $O/r02c7_bench_c3.json 2> $O/r02c7_bench_c3.err; echo "bench rc=$?"
MAU_FLAGS=16384 timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 1 > $O/r02c7_bench_c3_nostats.json 2> /dev/null; echo "bench no-conv-stats rc=$?"
timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 0 --profile-layers > /dev/null 2> $O/r02c7_layers_c3.txt; echo "layers rc=$?"
timeout 300 python tools/timeline.py --config 3 --tag r02c7 > $O/r02c7_timeline.txt 2>&1; echo "timeline rc=$?"
grep -A14 "step span" $O/r02c7_timeline.txt
python -c "
import json
for f in ('r02c7_bench_c3','r02c7_bench_c3_nostats'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
"
grep "k:conv.*fwd" $O/r02c7_layers_c3.txt | head -20
