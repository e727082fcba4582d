# final 1-GPU call of round 2: parity suite, the default bench line + the CPU reference arm, smoke, ncu evidence
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02f_pytest.log
python __graft_entry__.py smoke > $O/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02f_smoke.log
timeout 600 python bench.py > $O/r02f_bench_default.json 2> $O/r02f_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02f_bench_ref.json 2> /dev/null; echo "ref rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02f_bench_default.json'))
print(d['metric'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d['e2e'], 'sust', d['sustained']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['burst']['frac'], 'cpu', d['cpu_baseline']['value'])
print('inference', round(d['inference']['value'],1), d['inference']['e2e'])
for k,v in d['riders'].items(): print(k, round(v['value'],1))
"
bash tools/run_profiles_r02.sh
