python -c "
import iqb200
from iqb200 import api
print('scalar TFMA/s', api.fma_peak(0), 'packed f32x2 TFMA/s', api.fma_peak(0, packed=True))
"
