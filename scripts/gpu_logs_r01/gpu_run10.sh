timeout 600 python -m pytest tests -x -q -m gpu -k "tau_model" 2>&1 | grep -E "assert|Error|passed|failed" | head -12
python scripts/exp_threads.py 4 8 12 14
OMP_WAIT_POLICY=passive python scripts/exp_threads.py 8 14
