# 8-GPU confirmation run: H2D concurrency at 1/2/4/8 ranks, data-parallel parity, training bench (the driver's form), timeline
O=gpurun_out
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29621 "${@:2}"; }
for n in 1 2 4 8; do timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n tools/h2d_diag.py 2>/dev/null | tail -1 >> $O/r02n8_h2d_diag.jsonl; done
cat $O/r02n8_h2d_diag.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 tools/dp_parity.py > $O/r02n8_dp_parity.txt 2>&1; grep dp_parity $O/r02n8_dp_parity.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02n8_bench_default.json 2> $O/r02n8_bench_default.err; echo "bench n8 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus 8 --config 3 --no-cpu-baseline --sustain-s 1 --comm-ctas 16 --no-e2e > $O/r02n8_bench_c3_ctas16.json 2> /dev/null; echo "bench n8 ctas16 rc=$?"
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29625 tools/timeline.py --config 3 --tag r02n8 --comm-ctas 8 > $O/r02n8_timeline.txt 2>&1; echo "timeline rc=$?"
grep -A14 "step span" $O/r02n8_timeline.txt
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r02n8_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d.get('e2e') and round(d['e2e']['value'],1), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'],3))
        if 'inference' in d: print('   inference', round(d['inference']['value'],1), 'e2e', round(d['inference']['e2e']['value'],1), {k: round(v['value'],1) for k,v in d['riders'].items()})
    except Exception as e: print(f, 'ERR', e)
"
