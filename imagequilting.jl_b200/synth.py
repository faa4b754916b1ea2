"""Seeded synthetic inputs of the five BASELINE.json configs: shapes, tile sizes
and conditioning data.  Everything is generated with numpy.random.default_rng(1000 + config)."""
from __future__ import annotations

import numpy as np


def gaussian_field(shape, ranges, seed):
    """White noise -> Gaussian low-pass (per-axis range in voxels) -> mean 0 / variance 1, FP32."""
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal(shape)
    spec = np.fft.rfftn(noise)
    for ax, (n, r) in enumerate(zip(shape, ranges)):
        f = np.fft.rfftfreq(n) if ax == len(shape) - 1 else np.fft.fftfreq(n)
        g = np.exp(-2.0 * (np.pi * f * (r / 3.0)) ** 2)
        sh = [1] * len(shape)
        sh[ax] = g.size
        spec = spec * g.reshape(sh)
    field = np.fft.irfftn(spec, s=shape, axes=tuple(range(len(shape))))
    field = (field - field.mean()) / field.std()
    return np.asfortranarray(field.astype(np.float32))


def box_mean(a, win):
    from numpy.lib.stride_tricks import sliding_window_view
    pad = [(w // 2, w // 2) for w in win]
    ap = np.pad(a.astype(np.float64), pad, mode="edge")
    return np.asfortranarray(sliding_window_view(ap, win).mean(axis=tuple(range(-len(win), 0))).astype(np.float32))


def config(k, scale=1.0):
    """Returns dict(name, trainimg, tilesize, kwargs) for config k in 1..5.  `scale` < 1 shrinks the
    training image (tests); the tile size is kept."""
    seed = 1000 + k

    def sz(*dims):
        return tuple(max(int(round(d * scale)), 1) for d in dims)

    if k == 1:
        shape = sz(100, 100)
        f = gaussian_field(shape, (20, 4), seed)
        ti = np.asfortranarray((f > np.percentile(f, 70)).astype(np.float32))
        return dict(name="cfg1 2D binary channels 100x100 tile 30x30", trainimg=ti, tilesize=(30, 30),
                    kwargs=dict(overlap=(1 / 6, 1 / 6), nreal=1, path="raster", tol=0.1))
    if k == 2:
        shape = sz(512, 512)
        ti = gaussian_field(shape, (12, 12), seed)
        return dict(name="cfg2 2D Gaussian 512x512 tile 48x48 nreal=16", trainimg=ti, tilesize=(48, 48),
                    kwargs=dict(nreal=16))
    if k == 3:
        shape = sz(100, 100, 50)
        f = gaussian_field(shape, (10, 10, 4), seed)
        q1, q2 = np.percentile(f, [33.3, 66.6])
        ti = np.asfortranarray(((f > q1).astype(np.float32) + (f > q2).astype(np.float32)))
        g = gaussian_field(shape, (10, 10, 4), seed + 1)
        obs = ((g > q1).astype(np.float32) + (g > q2).astype(np.float32))
        rng = np.random.default_rng(seed)
        flat = rng.choice(int(np.prod(shape)), size=min(200, int(np.prod(shape)) // 50), replace=False)
        hard = {tuple(int(c) for c in np.unravel_index(i, shape, order="F")): float(obs[np.unravel_index(i, shape, order="F")])
                for i in flat}
        return dict(name="cfg3 3D facies 100x100x50 tile 20x20x10 hard data nreal=8", trainimg=ti,
                    tilesize=(20, 20, 10), kwargs=dict(hard=hard, nreal=8))
    if k == 4:
        shape = sz(200, 200, 80)
        ti = gaussian_field(shape, (15, 15, 5), seed)
        auxti = box_mean(ti, (9, 9, 3))
        aux = box_mean(gaussian_field(shape, (15, 15, 5), seed + 1), (9, 9, 3))
        return dict(name="cfg4 3D Gaussian 200x200x80 tile 30x30x12 soft data nreal=8", trainimg=ti,
                    tilesize=(30, 30, 12), kwargs=dict(soft=[(aux, auxti)], nreal=8))
    if k == 5:
        shape = sz(250, 250, 100)
        ti = gaussian_field(shape, (20, 20, 6), seed)
        return dict(name="cfg5 3D Gaussian 250x250x100 tile 40x40x16 nreal=64", trainimg=ti, tilesize=(40, 40, 16),
                    kwargs=dict(nreal=64))
    raise ValueError(k)
