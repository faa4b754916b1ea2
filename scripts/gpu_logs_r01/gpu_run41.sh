timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -1
for v in mb5 mb6 xf5; do
  echo "== fft variant $v"
  IQB200_LIB=$PWD/imagequilting.jl_b200/build/libiqb200_$v.so timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -1
done
