timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r01_resident.json 2> gpurun_out/bench_r01_resident.err; tail -c 200 gpurun_out/bench_r01_resident.json; tail -2 gpurun_out/bench_r01_resident.err
