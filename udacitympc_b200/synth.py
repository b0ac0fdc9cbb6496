"""Synthetic workloads of BASELINE.json / SURVEY.md 8(d): identical bits for the oracle and the GPU path.

RNG: std::mt19937_64 (implemented here in numpy so it is available without a C++ toolchain), seed 20261018, one
stream, problems in index order, U = (rng() >> 11) * 2^-53.  Host-side input generation only; nothing here is on
the measured path.
"""
import os

import numpy as np

SEED = 20261018
_PKG = os.path.dirname(os.path.abspath(__file__))


class MT19937_64:
    NN, MM = 312, 156
    UM = np.uint64(0xFFFFFFFF80000000)
    LM = np.uint64(0x7FFFFFFF)
    MAT = np.uint64(0xB5026F5AA96619E9)

    def __init__(self, seed=5489):
        mt = np.zeros(self.NN, dtype=np.uint64)
        x = seed & 0xFFFFFFFFFFFFFFFF
        mt[0] = x
        for i in range(1, self.NN):
            x = (6364136223846793005 * (x ^ (x >> 62)) + i) & 0xFFFFFFFFFFFFFFFF
            mt[i] = x
        self.mt = mt
        self.buf = np.zeros(0, dtype=np.uint64)

    def _twist(self):
        mt, NN, MM = self.mt, self.NN, self.MM

        def mix(hi, lo, far):
            x = (hi & self.UM) | (lo & self.LM)
            return far ^ (x >> np.uint64(1)) ^ np.where(x & np.uint64(1), self.MAT, np.uint64(0))

        new = mt.copy()
        new[:NN - MM] = mix(mt[:NN - MM], mt[1:NN - MM + 1], mt[MM:NN])              # i = 0..155
        new[NN - MM:NN - 1] = mix(mt[NN - MM:NN - 1], mt[NN - MM + 1:NN], new[:MM - 1])  # i = 156..310
        new[NN - 1] = mix(mt[NN - 1:NN], new[0:1], new[MM - 1:MM])[0]                # i = 311
        self.mt = new
        y = new.copy()
        y ^= (y >> np.uint64(29)) & np.uint64(0x5555555555555555)
        y ^= (y << np.uint64(17)) & np.uint64(0x71D67FFFEDA60000)
        y ^= (y << np.uint64(37)) & np.uint64(0xFFF7EEE000000000)
        y ^= y >> np.uint64(43)
        return y

    def raw(self, n):
        out = [self.buf]
        have = self.buf.size
        while have < n:
            blk = self._twist()
            out.append(blk)
            have += blk.size
        allv = np.concatenate(out)
        self.buf = allv[n:]
        return allv[:n]

    def uniform(self, n):
        return (self.raw(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def roadmap_centerline():
    """Centre-line (x, y) of the reference's mpc_to_line/roadmap.csv (columns 4, 5), 246 points.  Stored as a data
    fixture (udacitympc_b200/data/roadmap_centerline.csv, written by tests/golden/make_golden.py)."""
    return np.loadtxt(os.path.join(_PKG, "data", "roadmap_centerline.csv"), delimiter=",")


def line_problems(B, rng=None):
    """degree-1 reference y = -1 (coeffs of config 1), randomized x0/cte/epsi."""
    rng = rng or MT19937_64(SEED)
    u = rng.uniform(4 * B).reshape(B, 4)
    c0, c1 = -1.0, 0.0
    x = -10.0 + 20.0 * u[:, 0]
    fx = c0 + c1 * x
    y = fx - 8.0 + 16.0 * u[:, 1]
    psi = -0.6 + 1.2 * u[:, 2]
    v = 5.0 + 30.0 * u[:, 3]
    states = np.stack([x, y, psi, v, fx - y, psi - np.arctan(c1)], axis=1)
    coeffs = np.tile(np.array([c0, c1]), (B, 1))
    return np.ascontiguousarray(states), np.ascontiguousarray(coeffs)


def roadmap_windows(B, rng=None):
    """window w = b mod 240 of 6 consecutive centre-line points in the window-local
    frame (origin = first point, x-axis along the first segment), jitter 0.05*(U-0.5) m on each coordinate.
    Returns xs, ys of shape (B, 6)."""
    rng = rng or MT19937_64(SEED)
    cl = roadmap_centerline()
    w = np.arange(B) % 240
    idx = w[:, None] + np.arange(6)[None, :]
    px, py = cl[idx, 0], cl[idx, 1]
    ang = np.arctan2(py[:, 1] - py[:, 0], px[:, 1] - px[:, 0])
    dx, dy = px - px[:, :1], py - py[:, :1]
    ca, sa = np.cos(ang)[:, None], np.sin(ang)[:, None]
    lx = ca * dx + sa * dy
    ly = -sa * dx + ca * dy
    u = rng.uniform(12 * B).reshape(B, 12)
    lx = lx + 0.05 * (u[:, :6] - 0.5)
    ly = ly + 0.05 * (u[:, 6:] - 0.5)
    return np.ascontiguousarray(lx), np.ascontiguousarray(ly)


def roadmap_problems(B, coeffs, rng=None):
    """degree-3 coeffs (from the window fits, local frame), vehicle at the local origin."""
    rng = rng or MT19937_64(SEED + 1)
    u = rng.uniform(3 * B).reshape(B, 3)
    c0, c1 = coeffs[:, 0], coeffs[:, 1]
    x = np.zeros(B)
    y = -2.0 + 4.0 * u[:, 0]
    psi = np.arctan(c1) - 0.3 + 0.6 * u[:, 1]
    v = 5.0 + 30.0 * u[:, 2]
    states = np.stack([x, y, psi, v, c0 - y, psi - np.arctan(c1)], axis=1)
    return np.ascontiguousarray(states)


def kinematic_inputs(B, H=1, rng=None):
    """x,y~U(-100,100), psi~U(-pi,pi), v~U(0,40), delta~U(+-0.436332), a~U(-1,1)."""
    rng = rng or MT19937_64(SEED + 2)
    u = rng.uniform(4 * B).reshape(B, 4)
    states = np.stack([-100 + 200 * u[:, 0], -100 + 200 * u[:, 1], -np.pi + 2 * np.pi * u[:, 2], 40 * u[:, 3]], axis=1)
    a = rng.uniform(2 * B * H).reshape(B, H, 2)
    act = np.stack([-0.436332 + 2 * 0.436332 * a[..., 0], -1 + 2 * a[..., 1]], axis=2)
    return np.ascontiguousarray(states), np.ascontiguousarray(act)
