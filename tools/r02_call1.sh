# (historical: MAU_FLAGS=4096 selected the cooperative one-launch BatchNorm of that commit; the path was measured slower and removed -- profiles/r02_training_step.md)
# round 2, GPU call 1: parity suite, the new bench line (both arms), SSIM criterion, per-op profile of a training step,
# fused (cooperative) BatchNorm A/B
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02c1_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/r02c1_pytest.log
timeout 600 python bench.py > $O/r02c1_bench_default.json 2> $O/r02c1_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02c1_bench_ref.json 2> $O/r02c1_bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --config 3 --criterion l1-gradient-ssim --no-cpu-baseline --sustain-s 0 > $O/r02c1_bench_c3_ssim.json 2> $O/r02c1_bench_c3_ssim.err; echo "ssim rc=$?"
timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 0 --profile-layers > $O/r02c1_bench_c3.json 2> $O/r02c1_layers_c3.txt; echo "layers rc=$?"
MAU_FLAGS=4096 timeout 300 python bench.py --config 3 --no-cpu-baseline --sustain-s 0 --profile-layers > $O/r02c1_bench_c3_bnunfused.json 2> $O/r02c1_layers_c3_bnunfused.txt; echo "unfused rc=$?"
python __graft_entry__.py smoke > $O/r02c1_smoke.log 2>&1; echo "smoke rc=$?"
python -c "
import json
for f in ('r02c1_bench_default','r02c1_bench_c3','r02c1_bench_c3_bnunfused','r02c1_bench_c3_ssim'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['metric'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3))
    except Exception as e: print(f, 'ERR', e)
"
