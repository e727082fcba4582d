# 2-GPU development run: data-parallel parity (fp32 / bf16 wire), training bench variants, kernel timeline with NCCL
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tools/dp_parity.py > $O/r02n2_dp_parity_fp32.txt 2>&1; echo "dp_parity fp32 rc=$?"; grep dp_parity $O/r02n2_dp_parity_fp32.txt
timeout 300 $TR tools/dp_parity.py --grad-dtype=bf16 > $O/r02n2_dp_parity_bf16.txt 2>&1; echo "dp_parity bf16 rc=$?"; grep dp_parity $O/r02n2_dp_parity_bf16.txt
for v in "fp32 4" "bf16 4" "bf16 2" "fp32 8"; do set -- $v
  timeout 300 $TR bench.py --gpus 2 --config 3 --no-cpu-baseline --sustain-s 1 --grad-dtype $1 --comm-ctas $2 > $O/r02n2_bench_c3_$1_ctas$2.json 2> $O/r02n2_bench.err; echo "bench $1 ctas $2 rc=$?"
done
NCCL_DEBUG=INFO timeout 300 $TR tools/timeline.py --config 3 --tag r02n2 --comm-ctas 4 > $O/r02n2_timeline.txt 2>&1; echo "timeline rc=$?"
grep -E "step span|launches|idle gaps" $O/r02n2_timeline.txt | head -20
grep -E "NVLS|Ring|Tree|Channel|Proto|algo" $O/r02n2_timeline.txt | head -10
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r02n2_bench_c3_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'sust', d['sustained'] and round(d['sustained']['ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
"
