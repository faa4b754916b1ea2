echo "== default (in-place FFT, zdirect pair-fastest)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -x -q 2>&1 | tail -3
timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -1
for v in t512 t256 o3 o3t512 o3r4 o3r16; do
  echo "== cut variant $v"
  IQB200_LIB=$PWD/imagequilting.jl_b200/build/libiqb200_$v.so timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -1
done
IQB200_LIB=$PWD/imagequilting.jl_b200/build/libiqb200_o3.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -x -q -k "cut or resident" 2>&1 | tail -3
