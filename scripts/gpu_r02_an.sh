# round 2, call AN: z pass with packed FMAs (k_fft_zdirect2, IQB200_FFT_ZPACKED=1) against the scalar kernel on config 5,
# then the FFT-path parity tests with the packed kernel
run() { # zpacked
  IQB200_FFT_ZPACKED=$1 timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('zpacked $1: value %.1fM e2e %.1fM ms %.0f device %.0f cut %.1f dist %.1f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms']))"
}
run 0
run 1
IQB200_FFT_ZPACKED=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_parity.py -x -q -m gpu -k "fft or full_size or teacher or imfilter or resident" 2>&1 | tail -2
