O=gpurun_out
timeout 200 python tools/dp_diag1.py > $O/r02c8_dp_diag.txt 2>&1; grep -E "rep" $O/r02c8_dp_diag.txt
MAU_FLAGS=8192 timeout 200 python tools/dp_diag1.py > $O/r02c8_dp_diag_nooverlap.txt 2>&1; grep -E "rep" $O/r02c8_dp_diag_nooverlap.txt
timeout 600 python -m pytest tests -m gpu -q > $O/r02c8_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02c8_pytest.log
timeout 600 python bench.py > $O/r02c8_bench_default.json 2> $O/r02c8_bench_default.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02c8_bench_default.json'))
print(d['metric'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d['e2e'], 'sust', d['sustained']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['burst']['frac'])
print('inference', round(d['inference']['value'],1), d['inference']['e2e'])
for k,v in d['riders'].items(): print(k, round(v['value'],1), v.get('e2e'))
"
