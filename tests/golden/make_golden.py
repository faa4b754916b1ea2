#!/usr/bin/env python
"""Generates tests/golden/oracle_golden.npz from the CPU oracle (oracle/iq_oracle.py).

These are ORACLE outputs, not outputs of the Julia reference (which cannot run here):
they freeze the restatement so that an accidental change of the oracle -- the checker every GPU parity
test relies on -- is caught by the CPU suite.  Regenerate with:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import iq_oracle as O  # noqa: E402


def inputs():
    r = np.random.default_rng(2024)
    ti2 = r.integers(0, 3, (28, 24)).astype(np.float64)
    ti3 = r.standard_normal((16, 14, 8))
    kern = r.standard_normal((5, 4, 3))
    w = (r.random((5, 4, 3)) < 0.5).astype(np.float64)
    A, B = r.standard_normal((3, 9, 4)), r.standard_normal((3, 9, 4))
    D = r.integers(0, 6, 120).astype(np.float64)
    Da = r.integers(0, 9, 120).astype(np.float64)
    D[r.random(120) < 0.1] = np.inf
    Da[np.isinf(D)] = np.inf
    return dict(ti2=ti2, ti3=ti3, kern=kern, w=w, A=A, B=B, D=D, Da=Da)


def outputs(x):
    out = {}
    out["dist_direct"] = O.fastdistance(x["ti3"], x["kern"], x["w"], method="direct")
    out["dist_fft"] = O.fastdistance(x["ti3"], x["kern"], x["w"], method="fft")
    out["cut0"] = O.graphcut(x["A"], x["B"], 0)
    out["cut1"] = O.graphcut(x["A"], x["B"], 1)
    out["relax"] = O.relaxation(x["D"], [x["Da"]], 0.1)
    out["tau"] = O.taumodel(out["relax"], x["D"], [x["Da"]])
    hard = {(3, 4): 2.0, (20, 17): 0.0, (0, 0): float("nan")}
    aux = np.round(np.add.outer(np.arange(28), np.arange(24)) / 8.0)
    reals, cuts, voxs = O.iqsim(x["ti2"], (10, 8), (30, 26), overlap=(0.3, 0.25), soft=[(np.pad(aux, ((0, 2), (0, 2)), mode="edge"), aux)],
                                hard=hard, tol=0.2, path="dilation", nreal=2, debug=True, rng=np.random.default_rng(99))
    out["real0"], out["real1"] = reals
    out["cutgrid0"] = cuts[0]
    out["voxs"] = np.array(voxs)
    out["path"] = np.array(O.genpath(np.random.default_rng(5), (4, 3, 2), "dilation", []))
    return out


if __name__ == "__main__":
    x = inputs()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz"), **outputs(x))
    print("written")
