#!/bin/bash
# round 2, last call: BatchNorm backward kernels compiled for three resident blocks per SM (80 registers, some spills) and
# the head backward for four (64 registers) -- a second library built with -DMAU_BN_BWD_MIN_BLOCKS=3 -DMAU_HEAD_BWD_OCC4 --
# against the default build: per-kernel bandwidth, training step, then tests + smoke ON THE VARIANT.
# (the variant library is not part of the build: nvcc ... -DMAU_BN_BWD_MIN_BLOCKS=3 -DMAU_HEAD_BWD_OCC4 -c norm.cu / elementwise.cu,
# linked with the other objects of csrc/obj into metadata-augmented-unet-for-lst-ndvi_b200/libmau_b200_occ.so; it lost -- see profiles/r02_training_step.md)
O=gpurun_out; mkdir -p $O; P=metadata-augmented-unet-for-lst-ndvi_b200
export BW_ONLY=bn_bwd_reduce,bn_bwd_apply,head_bwd
B="python bench.py --config 3 --no-cpu-baseline --sustain-s 1 --no-e2e --no-riders"
timeout 100 python tools/bw_bench.py > $O/r02p_bw_default.txt 2>&1
timeout 200 $B > $O/r02p_c3_default.json 2> /dev/null; echo "default rc=$?"
cp $P/libmau_b200.so $P/libmau_b200_default.so.keep; cp $P/libmau_b200_occ.so $P/libmau_b200.so
timeout 100 python tools/bw_bench.py > $O/r02p_bw_occ.txt 2>&1
timeout 200 $B > $O/r02p_c3_occ.json 2> /dev/null; echo "variant rc=$?"
unset BW_ONLY
timeout 300 python -m pytest tests -m gpu -q -x > $O/r02p_pytest_occ.log 2>&1; echo "pytest (variant) rc=$?"; tail -2 $O/r02p_pytest_occ.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02p_smoke_occ.log 2>&1; echo "smoke (variant) rc=$?"
cp $P/libmau_b200_default.so.keep $P/libmau_b200.so
echo "-- default"; cat $O/r02p_bw_default.txt; echo "-- variant"; cat $O/r02p_bw_occ.txt
python -c "
import json
for f in ('default','occ'):
    d=json.load(open('gpurun_out/r02p_c3_%s.json' % f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), d['sustained']['ms_per_step'])"
