#!/bin/bash
# last sanity pass of round 2: GPU tests + smoke at HEAD, and the bandwidth table of the final kernels
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02s_smoke.log
timeout 200 python bench.py --config 3 --no-cpu-baseline --sustain-s 1 --no-e2e --no-riders > gpurun_out/r02s_c3.json 2> /dev/null; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02s_c3.json')); print(round(d['value'],1), round(d['ms_per_step'],3), d['sustained']['ms_per_step'])"
