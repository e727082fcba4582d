#!/bin/bash
# last sanity pass of round 2: GPU tests + smoke at HEAD
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02s_smoke.log
