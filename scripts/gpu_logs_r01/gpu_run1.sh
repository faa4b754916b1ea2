nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30
