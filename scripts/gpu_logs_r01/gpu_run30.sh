IQB200_LIB=$PWD/imagequilting.jl_b200/build/libiqb200_prof.so python scripts/cut_bench.py 2>&1 | grep -E "^cut nfree" | awk 'NR%6==1' | head -12
