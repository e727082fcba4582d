O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "train_step or unetpp or kat" > $O/pytest_gpu20.log 2>&1; echo pytest rc=$?; tail -2 $O/pytest_gpu20.log
for f in 0 2048; do
  MAU_FLAGS=$f timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --profile-layers 2> $O/layers_c4_f$f.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flags $f', d['config']['baseline_config'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['roofline']['frac'])"
done
paste <(grep -E "^conv0_[1-4].conv1.bwd|^conv1_[1-3].conv1.bwd|^emb.bwd|^conv2_[12].conv1.bwd|^conv3_1.conv1.bwd" $O/layers_c4_f0.txt) <(grep -E "^conv0_[1-4].conv1.bwd|^conv1_[1-3].conv1.bwd|^emb.bwd|^conv2_[12].conv1.bwd|^conv3_1.conv1.bwd" $O/layers_c4_f2048.txt | awk '{print $2}')
