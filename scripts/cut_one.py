import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iqb200
from iqb200 import api, synth
ti = synth.config(5)["trainimg"].astype(np.float64)
A = ti[10:17, 20:60, 30:46]; B = ti[100:107, 120:160, 50:66]
with api.SearchContext(np.zeros((16, 16), np.float32), (4, 4)) as ctx:
    for _ in range(3):
        k, it = ctx.cut_batch([(A, B, 0)])
    print(it)
