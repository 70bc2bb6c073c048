timeout 600 python -m pytest tests/test_gpu_layer.py -q -x --timeout 300 -k "other_dims" > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -n 25 gpurun_out/pytest.log
