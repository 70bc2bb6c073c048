timeout 600 python -m pytest tests/test_gpu_ep.py -q -x --timeout 200 > gpurun_out/pytest_ep.log 2>&1; echo "pytest_ep exit=$?"; tail -n 3 gpurun_out/pytest_ep.log
bash tools/_tmp_run.sh 2
