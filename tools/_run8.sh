timeout 1200 python -m pytest tests/test_gpu_cpp_path.py -x -q -m gpu > gpurun_out/r3a_pytest_cpp.log 2>&1; tail -15 gpurun_out/r3a_pytest_cpp.log
