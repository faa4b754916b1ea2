timeout 900 python -m pytest tests/test_gpu_resident.py -x -q 2>&1 | tail -12
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
