"""Host-side helpers of the drop-in API (reference: dabry/misc.py).

Only what the extremal-field path touches is provided: the coordinate / unit enums (misc.py:142-170), the scipy
event decorators (34-43), the dyadic site-naming helpers (46-69), the constrained-control helpers used on
obstacle boundaries (72-97), the N-linear clipped interpolation (570-589) for host-side field evaluation, the
triangle cost helper (115-139) and a wall-clock `Chrono` (592-618).  The device versions of the numerical helpers
live in dabry_b200/csrc; the ones here serve problem set-up (time upper bound, `in_obs`, ...) and the public
`FlowField.value` API.
"""
import csv
import math
import time
from enum import Enum

import numpy as np

DEG_TO_RAD = math.pi / 180.
RAD_TO_DEG = 180. / math.pi


def non_terminal(func):
    """Marks a scipy `solve_ivp` event as non terminal, triggered on downward zero crossings (misc.py:34-37)."""
    func.terminal = False
    func.direction = -1.
    return func


def terminal(func):
    """Marks a scipy `solve_ivp` event as terminal, triggered on downward zero crossings (misc.py:40-43)."""
    func.terminal = True
    func.direction = -1.
    return func


def to_alpha(i: int) -> str:
    """Base-26 spelling with letters A..Z, 0 -> 'A' (misc.py:46-57)."""
    if i == 0:
        return 'A'
    letters = []
    while i > 0:
        letters.append(chr(65 + i % 26))
        i //= 26
    return ''.join(reversed(letters))


def alpha_to_int(s: str) -> int:
    """Inverse of :func:`to_alpha` (misc.py:60-63)."""
    value = 0
    for ch in s:
        value = 26 * value + (ord(ch) - 65)
    return value


def diadic_valuation(i: int) -> int:
    """Number of trailing zero bits of a positive integer (misc.py:66-69)."""
    if i <= 0:
        raise ValueError('diadic_valuation needs a positive integer')
    return (i & -i).bit_length() - 1


def directional_timeopt_control(ff_val, d, srf_max: float):
    """Control of norm `srf_max` such that the ground velocity `ff_val + u` points along `d` (misc.py:72-84).

    With e = d/|d| and n = R(pi/2) e, the cross-track flow is c = ff_val . n; the control is -n rotated by
    atan2(sqrt(max(srf^2 - c^2, 0)), c)."""
    d = np.asarray(d, dtype=float)
    e = d / np.linalg.norm(d)
    n = np.array((-e[1], e[0]))
    cross_track = np.asarray(ff_val, dtype=float) @ n
    angle = np.arctan2(np.sqrt(np.max((srf_max ** 2 - cross_track ** 2, 0.))), cross_track)
    ca, sa = np.cos(angle), np.sin(angle)
    rot = np.array(((ca, -sa), (sa, ca)))
    return srf_max * rot @ (-n)


def is_possible_direction(ff_val, d, srf_max: float) -> bool:
    """True when the cross-track flow is weaker than `srf_max`, i.e. the ground velocity can align with `d`
    (misc.py:87-97)."""
    d = np.asarray(d, dtype=float)
    e = d / np.linalg.norm(d)
    n = np.array((-e[1], e[0]))
    return bool(np.abs(np.asarray(ff_val, dtype=float) @ n) < srf_max)


def csv_to_dict(csv_file_path):
    """First column -> {header: value} for every non-empty row (misc.py:100-112)."""
    table = {}
    with open(csv_file_path, 'r') as handle:
        reader = csv.reader(handle)
        header = next(reader)
        for row in reader:
            if not row or len(row[0]) == 0:
                continue
            table[row[0]] = dict(zip(header[1:], row[1:]))
    return table


def _cross2(a, b):
    return a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]


def triangle_mask_and_cost(x, p1, p2, p3, c1: float, c2: float, c3: float):
    """Barycentric interpolation of (c1, c2, c3) over triangle (p1, p2, p3) on the grid `x` (nx, ny, 2); +inf
    outside the open triangle and for degenerate triangles (misc.py:115-139)."""
    v1 = p2 - p1
    v2 = p3 - p1
    out = np.full(x.shape[:-1], np.inf)
    den = _cross2(v1, v2)
    if np.isclose(den, 0.):
        return out
    a = (_cross2(x, v2) - _cross2(p1, v2)) / den
    b = -(_cross2(x, v1) - _cross2(p1, v1)) / den
    inside = (a > 0) & (b > 0) & (a + b < 1)
    out[inside] = ((1 - (a + b)) * c1 + a * c2 + b * c3)[inside]
    return out


class Coords(Enum):
    CARTESIAN = 'cartesian'
    GCS = 'gcs'

    @classmethod
    def from_string(cls, s):
        s = str(s)
        for member in cls:
            if member.value == s:
                return member
        raise ValueError('Unknown coord type "%s"' % s)


class Units(Enum):
    METERS = 'meters'
    RADIANS = 'rad'
    DEGREES = 'degrees'

    @classmethod
    def from_string(cls, s):
        s = str(s)
        for member in cls:
            if member.value == s:
                return member
        raise ValueError('Unknown units "%s"' % s)


class Utils:
    DEG_TO_RAD = DEG_TO_RAD
    RAD_TO_DEG = RAD_TO_DEG
    AIRSPEED_DEFAULT = 23.  # [m/s] (misc.py:176)
    EARTH_RADIUS = 6378.137e3  # [m] equatorial radius (misc.py:183)

    @staticmethod
    def ang_principal(a):
        """Angle in radians brought to ]-pi, pi] (misc.py:214-221)."""
        return np.angle(np.cos(a) + 1j * np.sin(a))

    @staticmethod
    def to_0_360(angle):
        """Angle in degrees brought to [0, 360[ (misc.py:186-192)."""
        return angle - 360. * math.floor(angle / 360.)

    @staticmethod
    def rectify(a, b):
        """Two angles in degrees brought to [0, 720[ keeping their order (misc.py:200-211)."""
        lo, hi = Utils.to_0_360(a), Utils.to_0_360(b)
        return lo, hi + 360. if lo > hi else hi

    @staticmethod
    def angular_diff(a1, a2):
        """Unsigned angle between two directions in radians (misc.py:223-235)."""
        d = abs(Utils.ang_principal(a2) - Utils.ang_principal(a1))
        return d if d <= np.pi else 2 * np.pi - d

    @staticmethod
    def in_lonlat_box(bl, tr, p):
        """True when (lon, lat) = p in radians lies strictly inside the box [bl, tr] (misc.py:558-568)."""
        lon_lo, lon_hi = Utils.rectify(Utils.RAD_TO_DEG * bl[0], Utils.RAD_TO_DEG * tr[0])
        lon = Utils.to_0_360(Utils.RAD_TO_DEG * p[0])
        return (lon_lo < lon < lon_hi or lon_lo < lon + 360 < lon_hi) and bl[1] < p[1] < tr[1]

    @staticmethod
    def central_angle_cos(x1, x2):
        return math.sin(x1[1]) * math.sin(x2[1]) + math.cos(x1[1]) * math.cos(x2[1]) * math.cos(x2[0] - x1[0])

    @staticmethod
    def distance(x1, x2, coords: Coords):
        """Euclidean distance, or great-circle distance in metres for (lon, lat) radians (misc.py:365-374)."""
        if coords == Coords.GCS:
            return math.acos(min(1., max(-1., Utils.central_angle_cos(x1, x2)))) * Utils.EARTH_RADIUS
        return float(np.linalg.norm(np.asarray(x1) - np.asarray(x2)))

    @staticmethod
    def interpolate(values, bl, spacings, state_ex, ndim_values_data=1):
        """N-linear interpolation of node data with clipping to the border (misc.py:570-589).

        The high-side weight is the fractional part of the UNCLIPPED grid position; the two corner indices are
        clipped independently to [0, n - 1]; the corner weights are the products (w_0 * w_1) * w_2 in axis order."""
        state_ex = np.asarray(state_ex, dtype=float)
        nd = state_ex.shape[0]
        pos = (state_ex - np.asarray(bl)) / np.asarray(spacings)
        lo = np.floor(pos).astype(np.int32)
        w_hi = pos - lo
        w_lo = 1 - w_hi
        top = np.array(values.shape[:nd]) - 1
        i_lo = np.clip(lo, 0, top)
        i_hi = np.clip(lo + 1, 0, top)
        acc = 0.
        for corner in range(1 << nd):
            weight = 1.
            index = []
            for axis in range(nd):
                high = (corner >> (nd - 1 - axis)) & 1
                w = w_hi[axis] if high else w_lo[axis]
                weight = w if axis == 0 else weight * w
                index.append(i_hi[axis] if high else i_lo[axis])
            acc = acc + weight * values[tuple(index)]
        return acc

    @staticmethod
    def time_fmt(duration):
        """Human readable duration (used by Chrono)."""
        if duration < 200:
            return f'{duration:.2f}s'
        if duration < 60 * 200:
            return f'{duration / 60:.1f}min'
        return f'{duration / 3600:.1f}h'


class Chrono:
    """Wall-clock context manager (misc.py:592-618)."""

    def __init__(self, no_verbose=False):
        self.t_start = 0.
        self.t_end = 0.
        self.verbose = not no_verbose

    def __enter__(self):
        self.start()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.stop()

    def start(self, msg=''):
        if self.verbose:
            print(f'[*] {msg}')
        self.t_start = time.time()

    def stop(self):
        self.t_end = time.time()
        if self.verbose:
            print(f'[*] Done ({self})')
        return self.t_end - self.t_start

    def __str__(self):
        return Utils.time_fmt(self.t_end - self.t_start)
