# round 2, call I: measured direct-vs-FFT crossover sweep + FFMA / FFMA2 peak
python - <<'PY'
from iqb200 import api
print('scalar TFMA/s', api.fma_peak(0), 'packed f32x2 TFMA/s', api.fma_peak(0, packed=True))
PY
timeout 1200 python scripts/crossover_sweep.py > gpurun_out/r02_crossover.log 2>&1
tail -25 gpurun_out/r02_crossover.log
