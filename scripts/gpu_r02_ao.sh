# round 2, calls AO / AP (run inline): sanity of the rebuilt default library after comment-only / experiments-only changes
timeout 250 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
