"""Synthetic flow grids of the shapes BASELINE.json names.

Plain numpy: returns arrays and problem parameters so that the same inputs can be fed to this package's
``DiscreteFF``/``NavigationProblem`` or (in the container, for golden vectors) to the reference's classes.
The generators are deterministic (fixed ``default_rng`` seeds) and size-parametrised so that tests can use
small grids of exactly the same family as the benchmark grids.
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

DEG = np.pi / 180.
SRF_AIRSPEED = 23.  # m/s, reference default airspeed (misc.py:176)


@dataclass
class SyntheticCase:
    """values: (nt, nx, ny, 2) fp64; bounds: (3, 2) rows (t, x, y); the rest are NavigationProblem kwargs."""
    values: np.ndarray
    bounds: np.ndarray
    coords: str  # 'cartesian' | 'gcs'
    x_init: np.ndarray
    x_target: np.ndarray
    srf: float
    bl: np.ndarray
    tr: np.ndarray
    total_duration: Optional[float] = None
    circle_obstacles: list = field(default_factory=list)  # [(center(2,), radius)]
    circle_penalty: Optional[tuple] = None  # (center(2,), radius, amplitude)
    discrete_penalty: Optional[tuple] = None  # (data (nt, nx, ny), ts (nt,), grid (nx, ny, 2)) of a DiscretePenalty
    name: str = 'synthetic'


def planar_waves(nt: int = 48, nx: int = 1024, ny: int = 1024, seed: int = 20231019) -> SyntheticCase:
    """Config 3: planar unsteady field, a sum of four travelling divergence-free cells on [0,1]^2, t in [0,2]."""
    rng = np.random.default_rng(seed)
    ks = 2 * np.pi * np.array((1., 2., 3., 5.))
    ls = 2 * np.pi * np.array((2., 1., 5., 3.))
    oms = np.array((0.5, 1., 1.5, 2.))
    alphas = rng.uniform(0, 2 * np.pi, 4)
    betas = rng.uniform(0, 2 * np.pi, 4)
    amps = np.array((1., 0.6, 0.35, 0.2))
    t = np.linspace(0., 2., nt)[:, None, None]
    x = np.linspace(0., 1., nx)[None, :, None]
    y = np.linspace(0., 1., ny)[None, None, :]
    u = np.zeros((nt, nx, ny))
    v = np.zeros((nt, nx, ny))
    for k, l, om, a, b, amp in zip(ks, ls, oms, alphas, betas, amps):
        ph = k * x + om * t + a
        u += -amp * np.sin(ph) * np.cos(l * y + b)
        v += amp * np.cos(ph) * np.sin(l * y + b)
    values = np.stack((u, v), -1)
    srf = 1.
    values *= 0.8 * srf / np.max(np.linalg.norm(values, axis=-1))
    bounds = np.array(((0., 2.), (0., 1.), (0., 1.)))
    return SyntheticCase(values, bounds, 'cartesian', np.array((0.2, 0.5)), np.array((0.8, 0.5)), srf,
                         np.array((0., 0.)), np.array((1., 1.)), total_duration=1.5, name='planar_waves')


def _smooth5(a: np.ndarray, axis: int) -> np.ndarray:
    """5-node box smoothing along one axis with edge replication."""
    pad = [(0, 0)] * a.ndim
    pad[axis] = (2, 2)
    ap = np.pad(a, pad, mode='edge')
    n = a.shape[axis]
    acc = np.zeros_like(a)
    for s in range(5):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(s, s + n)
        acc += ap[tuple(sl)]
    return acc / 5.


def _wind_on_sphere(nt, lon, lat, t_span, seed, sigma=2.):
    rng = np.random.default_rng(seed)
    t_w = 24 * 3600.
    t = np.linspace(0., t_span, nt)[:, None, None]
    lam = lon[None, :, None]
    phi = lat[None, None, :]
    u = 20. * np.cos(phi) * (1. + 0.5 * np.sin(3 * lam + 2 * np.pi * t / t_w))
    v = 8. * np.sin(2 * lam) * np.cos(phi) * np.cos(2 * np.pi * t / t_w)
    noise = rng.normal(0., 1., (2, nt, lon.shape[0], lat.shape[0]))
    for ax in (2, 3):
        noise = _smooth5(noise, ax)
    noise *= sigma / noise.std()
    return np.stack((u + noise[0], v + noise[1]), -1)


def era5_like(nt: int = 24, nx: int = 1440, ny: int = 721, seed: int = 20231020) -> SyntheticCase:
    """Config 4: ERA5-shaped global grid (0.25 deg when 1440x721), hourly frames, winds in m/s, GCS radians."""
    lon = np.linspace(-np.pi, np.pi - 2 * np.pi / nx, nx)
    lat = np.linspace(-np.pi / 2, np.pi / 2, ny)
    t_span = (nt - 1) * 3600. * (24. / nt if nt != 24 else 1.)
    values = _wind_on_sphere(nt, lon, lat, t_span, seed)
    bounds = np.array(((0., t_span), (lon[0], lon[-1]), (lat[0], lat[-1])))
    return SyntheticCase(values, bounds, 'gcs', DEG * np.array((-60., 40.)), DEG * np.array((-10., 50.)),
                         SRF_AIRSPEED, DEG * np.array((-80., 20.)), DEG * np.array((10., 70.)), name='era5_like')


def regional_constrained(nt: int = 17, nx: int = 181, ny: int = 101, seed: int = 20231021) -> SyntheticCase:
    """Config 5: what ``NavigationProblem.from_database`` builds (problem.py:298-307): a 0.5 deg regional box,
    3-hourly frames over 48 h, plus one circular obstacle and one circular penalty."""
    bl = DEG * np.array((-80., 20.))
    tr = DEG * np.array((10., 70.))
    lon = np.linspace(bl[0], tr[0], nx)
    lat = np.linspace(bl[1], tr[1], ny)
    t_span = 48 * 3600.
    values = _wind_on_sphere(nt, lon, lat, t_span, seed)
    bounds = np.array(((0., t_span), (bl[0], tr[0]), (bl[1], tr[1])))
    x_init = DEG * np.array((-60., 40.))
    x_target = DEG * np.array((-10., 50.))
    mid = 0.5 * (x_init + x_target)
    chord = x_target - x_init
    normal = np.array((-chord[1], chord[0])) / np.linalg.norm(chord)
    ext = np.min(tr - bl)
    obs = [(mid + 0.02 * np.linalg.norm(chord) * normal, 0.06 * ext)]
    pen = (x_init + 0.3 * chord - 0.05 * np.linalg.norm(chord) * normal, 0.1 * ext, 1.)
    return SyntheticCase(values, bounds, 'gcs', x_init, x_target, SRF_AIRSPEED, bl, tr,
                         circle_obstacles=obs, circle_penalty=pen, name='regional_constrained')


def regional_discrete_penalty(nt: int = 9, nx: int = 91, ny: int = 51, seed: int = 20231021) -> SyntheticCase:
    """Config 5 variant for SURVEY 8(f) rank 4: the regional case with its circular obstacle, the penalty given as a
    `DiscretePenalty` - a pulsating Gaussian bump sampled on a coarse (7, 31, 21) grid over the box and the 48 h."""
    sc = regional_constrained(nt, nx, ny, seed)
    ts = np.linspace(0., 48 * 3600., 7)
    xs = np.linspace(sc.bl[0], sc.tr[0], 31)
    ys = np.linspace(sc.bl[1], sc.tr[1], 21)
    chord = sc.x_target - sc.x_init
    normal = np.array((-chord[1], chord[0])) / np.linalg.norm(chord)
    c = sc.x_init + 0.3 * chord - 0.05 * np.linalg.norm(chord) * normal
    gx, gy = np.meshgrid(xs, ys, indexing='ij')
    bump = 4e-3 * np.exp(-((gx - c[0]) ** 2 + (gy - c[1]) ** 2) / (2 * 0.07 ** 2))
    data = bump[None, :, :] * (1. + 0.3 * np.sin(2 * np.pi * ts / (24 * 3600.)))[:, None, None]
    sc.circle_penalty = None
    sc.discrete_penalty = (data, ts, np.stack((gx, gy), -1))
    sc.name = 'regional_discrete_penalty'
    return sc
