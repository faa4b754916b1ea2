import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, iqb200
from iqb200 import synth
cfg = synth.config(1)
a = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=5, rng=np.random.default_rng(6), pipeline="staged", cut="host")
b = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=5, rng=np.random.default_rng(6), pipeline="staged", cut="device")
c = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=5, rng=np.random.default_rng(6), pipeline="resident")
print("host vs device cut (staged):", [bool(np.array_equal(x, y)) for x, y in zip(a, b)])
print("device-cut staged vs resident:", [bool(np.array_equal(x, y)) for x, y in zip(b, c)])
