O=gpurun_out
timeout 300 python tools/timeline.py --config 3 --tag r02c5 > $O/r02c5_timeline.txt 2>&1; echo "timeline rc=$?"
tail -20 $O/r02c5_timeline.txt
timeout 300 python -m pytest tests/test_zz_ssim_gpu.py -m gpu -q 2>&1 | tail -3
timeout 120 python tools/ssim_gpu_check.py 2>/dev/null | tail -1 > $O/r02c5_ssim_check.json; head -c 1500 $O/r02c5_ssim_check.json
