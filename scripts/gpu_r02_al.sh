# round 2, call AL: the reference-property tests with the IQ wrapper test, smoke
timeout 200 python -m pytest tests/test_gpu_reference_properties.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
