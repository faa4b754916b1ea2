"""Small workloads covering every kernel of the resident and host-staged pipelines (for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import iqb200
from iqb200 import synth

ti3 = synth.gaussian_field((40, 36, 18), (5, 5, 3), 9)
other = synth.gaussian_field((40, 36, 18), (5, 5, 3), 12)
auxti = np.asfortranarray(synth.box_mean(ti3, (5, 5, 3)).astype(np.float32))
aux = np.asfortranarray(synth.box_mean(other, (5, 5, 3)).astype(np.float32))
r = np.random.default_rng(1)
flat = r.choice(other.size, size=12, replace=False)
hard = {tuple(int(v) for v in c): float(other[tuple(c)]) for c in np.array(np.unravel_index(flat, other.shape)).T}
for pipeline in ("resident", "staged"):
    for fft in (-1, 1):
        iqb200.iqsim(ti3, (14, 12, 8), nreal=3, rng=np.random.default_rng(2), pipeline=pipeline, fft=fft, overlap=(0.25, 0.25, 0.25))
    iqb200.iqsim(ti3, (14, 12, 8), nreal=2, rng=np.random.default_rng(3), pipeline=pipeline, overlap=(0.25, 0.25, 0.25), soft=[(aux, auxti)])
    iqb200.iqsim(ti3, (14, 12, 8), nreal=2, rng=np.random.default_rng(4), pipeline=pipeline, overlap=(0.25, 0.25, 0.25), hard=hard, debug=True)
ti2 = synth.gaussian_field((72, 64), (5, 5), 3)
iqb200.iqsim(ti2, (20, 16), nreal=3, rng=np.random.default_rng(5), pipeline="resident", path="random", debug=True)
print("sanitize workload done")
