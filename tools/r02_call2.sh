# (historical: MAU_FLAGS=4096 selected the cooperative one-launch BatchNorm of that commit; the path was measured slower and removed -- profiles/r02_training_step.md)
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02c2_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $O/r02c2_pytest.log
MAU_FLAGS=4096 timeout 600 python -m pytest tests/test_bf16_layers_gpu.py tests/test_gpu_parity.py -m gpu -q -k "teacher or train_step or full_size_training" > $O/r02c2_pytest_bnfused.log 2>&1; echo "pytest bnfused rc=$?"
tail -5 $O/r02c2_pytest_bnfused.log
