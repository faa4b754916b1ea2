# round 2, call A: whole GPU suite (incl. the new full-size / teacher-forced parity tests) + baseline bench line
nproc; nvidia-smi --query-gpu=name --format=csv,noheader | head -1
timeout 2400 python -m pytest tests -q -m gpu --durations=15 2>&1 | tail -40 > gpurun_out/r02_a_tests.log
tail -5 gpurun_out/r02_a_tests.log


