timeout 600 python -m pytest tests/test_gpu_layer.py -q -x --timeout 300 -k "tf32" > gpurun_out/pytest.log 2>&1; echo "pytest tf32 exit=$?"; tail -n 25 gpurun_out/pytest.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_all.log 2>&1; echo "pytest all exit=$?"; tail -n 4 gpurun_out/pytest_all.log
