timeout 300 python tools/lstm_diag.py 2>&1 | grep "^B="
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for i in 1 2; do timeout 300 $TR tools/dp_parity.py 2>&1 | grep dp_parity; done
