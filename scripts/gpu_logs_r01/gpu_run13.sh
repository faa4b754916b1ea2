python scripts/exp_threads.py 14
