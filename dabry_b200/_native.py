"""ctypes binding of libdabry_b200.so (include/dabry_b200.h).

The library is the only execution path of the solver: if it has not been built (`python -c "import
__graft_entry__ as g; g.build()"`) or no sm_100 device is present, solver construction raises - there is no CPU
fallback.  `use_library(path)` exists for the CPU-only test tier, which loads tests/hostsim/libdabry_hostsim.so
(the same device headers compiled by g++) to exercise the host logic without a GPU; nothing in this package calls it.
"""
import ctypes as C
import time
import os

import numpy as np

ABI_VERSION = 2
COMM_ID_BYTES = 128
MAX_TARGET_EVENTS = 4

SIMPLE, RESAMPLING, TRIMMING = 0, 1, 2
CARTESIAN, GCS = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
# DABRY_B200_LIBRARY: another build of the CUDA library (tuning experiments); there is still no CPU path
DEFAULT_LIBRARY = os.environ.get('DABRY_B200_LIBRARY', os.path.join(_HERE, 'libdabry_b200.so'))

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class Config(C.Structure):
    _fields_ = [
        ('abi_version', C.c_int32), ('coords', C.c_int32), ('n_time', C.c_int32), ('max_depth', C.c_int32),
        ('mode_origin', C.c_int32), ('integrator', C.c_int32), ('dtype', C.c_int32), ('device', C.c_int32),
        ('n_index_per_subframe', C.c_int32), ('trimming_band_width', C.c_int32),
        ('cost_map_nx', C.c_int32), ('cost_map_ny', C.c_int32),
        ('n_roots', C.c_int64), ('root_offset', C.c_int64), ('root_count', C.c_int64), ('max_sites', C.c_int64),
        ('shard_block', C.c_int64), ('shard_world', C.c_int64),
        ('t_init', C.c_double), ('max_step', C.c_double), ('rtol', C.c_double), ('atol', C.c_double),
        ('srf', C.c_double), ('x_init', C.c_double * 2), ('x_target', C.c_double * 2),
        ('target_radius', C.c_double), ('max_dist', C.c_double), ('bl', C.c_double * 2), ('tr', C.c_double * 2),
        ('ff_max_norm', C.c_double), ('times', c_double_p),
    ]


class Wrapper(C.Structure):
    _fields_ = [('time_origin', C.c_double), ('scale_time', C.c_double), ('scale_length', C.c_double),
                ('bl', C.c_double * 2), ('scaler_speed', C.c_double), ('scaler_dspeed', C.c_double)]


class Leaf(C.Structure):
    _fields_ = [('kind', C.c_int32), ('nt', C.c_int32), ('table_off', C.c_int32), ('pad_', C.c_int32),
                ('coeff', C.c_double), ('t_start', C.c_double), ('t_end', C.c_double), ('p', C.c_double * 8)]


class Obstacle(C.Structure):
    _fields_ = [('kind', C.c_int32), ('pad_', C.c_int32), ('scale_length', C.c_double), ('bl', C.c_double * 2),
                ('p', C.c_double * 8)]


class Penalty(C.Structure):
    _fields_ = [('kind', C.c_int32), ('pad_', C.c_int32), ('time_origin', C.c_double), ('scale_time', C.c_double),
                ('scale_length', C.c_double), ('bl', C.c_double * 2), ('scaler', C.c_double), ('dx', C.c_double),
                ('p', C.c_double * 4)]


SITE_RECORD_DTYPE = np.dtype([
    ('dyadic', np.int64), ('index_t_init', np.int32), ('n_samples', np.int32), ('closure', np.int32),
    ('neuter', np.int32), ('index_neutered', np.int32), ('obstacle', np.int32), ('index_t_obs', np.int32),
    ('obs_trigo', np.int32), ('index_t_check_next', np.int32), ('depth', np.int32), ('parent_prev', np.int32),
    ('parent_next', np.int32), ('n_target_events', np.int32), ('ivp_status', np.int32), ('flags', np.int32), ('pad_', np.int32),
    ('t_enter_obs', np.float64), ('cost0', np.float64), ('t_target_events', np.float64, (MAX_TARGET_EVENTS,)),
], align=True)


class Stats(C.Structure):
    _fields_ = [('rk_steps', C.c_uint64), ('rhs_evals', C.c_uint64), ('ivp_calls', C.c_uint64),
                ('kernel_launches', C.c_uint64), ('n_sites', C.c_int64), ('capacity', C.c_int64),
                ('rounds', C.c_int32), ('pad_', C.c_int32), ('ms_integrate', C.c_double),
                ('ms_resample', C.c_double), ('ms_trim', C.c_double), ('ms_arrival', C.c_double),
                ('ms_total', C.c_double), ('ms_upload', C.c_double), ('ms_obstacle', C.c_double),
                ('ms_comm', C.c_double), ('rk_steps_free', C.c_uint64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_ if name != 'pad_'}


# every symbol include/dabry_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SIGNATURES = {
    'dabry_abi_version': (C.c_int, []),
    'dabry_create': (C.c_int, [C.POINTER(Config), C.POINTER(_H)]),
    'dabry_destroy': (None, [_H]),
    'dabry_last_error': (C.c_char_p, [_H]),
    'dabry_backend_name': (C.c_char_p, []),
    'dabry_host_register': (C.c_int, [C.c_void_p, C.c_size_t]),
    'dabry_host_unregister': (C.c_int, [C.c_void_p]),
    'dabry_release_cached_memory': (C.c_int, [C.c_int32]),
    'dabry_set_flow_grid': (C.c_int, [_H, c_double_p, c_double_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      c_double_p, c_double_p, C.POINTER(Wrapper)]),
    'dabry_set_flow_program': (C.c_int, [_H, C.POINTER(Leaf), C.c_int32, c_double_p, C.c_int64, C.POINTER(Wrapper)]),
    'dabry_set_obstacles': (C.c_int, [_H, C.POINTER(Obstacle), C.c_int32]),
    'dabry_set_penalty': (C.c_int, [_H, C.POINTER(Penalty)]),
    'dabry_set_penalty_grid': (C.c_int, [_H, C.POINTER(Penalty), c_double_p, c_double_p, c_double_p, c_double_p, C.c_int32,
                                         C.c_int32, C.c_int32]),
    'dabry_setup': (C.c_int, [_H, C.c_int32, c_double_p]),
    'dabry_step': (C.c_int, [_H]),
    'dabry_step_integrate': (C.c_int, [_H]),
    'dabry_step_resample': (C.c_int, [_H]),
    'dabry_step_trim_integrate': (C.c_int, [_H]),
    'dabry_step_trim_refine': (C.c_int, [_H]),
    'dabry_step_trim_close': (C.c_int, [_H]),
    'dabry_set_cost_map': (C.c_int, [_H, c_double_p]),
    'dabry_finish': (C.c_int, [_H]),
    'dabry_solve': (C.c_int, [_H, C.c_int32, c_double_p]),
    'dabry_num_sites': (C.c_int, [_H, C.POINTER(C.c_int64)]),
    'dabry_get_sites': (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int64]),
    'dabry_get_trajs': (C.c_int, [_H, C.c_int64, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p,
                                  c_double_p]),
    'dabry_get_next_nb': (C.c_int, [_H, C.c_int64, C.c_int64, c_int32_p]),
    'dabry_get_arrivals': (C.c_int, [_H, C.c_int64, C.c_int64, c_int32_p, c_double_p]),
    'dabry_get_solution_sites': (C.c_int, [_H, C.c_double, c_int32_p, C.c_int64, C.POINTER(C.c_int64)]),
    'dabry_get_solution': (C.c_int, [_H, C.POINTER(C.c_int64), c_double_p]),
    'dabry_get_cost_map': (C.c_int, [_H, C.c_int32, c_double_p]),
    'dabry_get_stats': (C.c_int, [_H, C.POINTER(Stats)]),
    'dabry_get_flow_max_norm': (C.c_int, [_H, c_double_p]),
    'dabry_boundary_doubles': (C.c_int, [_H, C.POINTER(C.c_int64)]),
    'dabry_boundary_count': (C.c_int, [_H, C.POINTER(C.c_int64)]),
    'dabry_export_boundary': (C.c_int, [_H, c_double_p]),
    'dabry_import_boundary': (C.c_int, [_H, c_double_p]),
    'dabry_export_boundary_device': (C.c_int, [_H, C.c_void_p]),
    'dabry_import_boundary_device': (C.c_int, [_H, C.c_void_p]),
    'dabry_get_cost_map_device': (C.c_int, [_H, C.c_int32, C.c_void_p]),
    'dabry_set_cost_map_device': (C.c_int, [_H, C.c_void_p]),
    'dabry_comm_unique_id': (C.c_int, [C.c_void_p]),
    'dabry_comm_create': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    'dabry_comm_destroy': (None, [C.c_void_p]),
    'dabry_comm_last_error': (C.c_char_p, []),
    'dabry_attach_comm': (C.c_int, [_H, C.c_void_p]),
    'dabry_set_flow_grid_shared': (C.c_int, [_H, c_double_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_double_p,
                                             c_double_p, C.POINTER(Wrapper)]),
    'dabry_solve_sharded': (C.c_int, [_H, C.c_int32, c_double_p]),
    'dabry_get_global_solution': (C.c_int, [_H, c_double_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    'dabry_time_upper_bound': (C.c_int, [_H, c_double_p, C.c_double, C.c_double, c_double_p]),
    'dabry_eval_rhs': (C.c_int, [_H, C.c_int64, c_double_p, c_double_p, c_double_p]),
    'dabry_eval_flow': (C.c_int, [_H, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p]),
}

_lib = None
_lib_path = DEFAULT_LIBRARY


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'libdabry_b200 error {code}: {message}')
        self.code = code


def use_library(path):
    """Test hook: bind another build of the same C ABI (tests/hostsim).  Not used by the package."""
    global _lib, _lib_path
    _lib = None
    _lib_path = DEFAULT_LIBRARY if path is None else path


def library():
    """The bound shared library; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_lib_path):
            raise ImportError(
                f'{_lib_path} not found: build the CUDA extension first (python -c "import __graft_entry__ as g; '
                f'g.build()" at the repository root).  dabry_b200 has no CPU execution path.')
        lib = C.CDLL(_lib_path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.dabry_abi_version() != ABI_VERSION:
            raise ImportError('libdabry_b200 ABI version mismatch')
        _lib = lib
    return _lib


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def iptr(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


CALL_TIMES = None  # diagnostic: set to a dict to accumulate the wall time of every native call by entry point


def _now():
    return time.perf_counter() if CALL_TIMES is not None else 0.


def _note(name, t0):
    if CALL_TIMES is not None:
        CALL_TIMES[name] = CALL_TIMES.get(name, 0.) + time.perf_counter() - t0


def release_cached_memory(device=0):
    """Give the device blocks parked by closed solvers back to the driver (dabry_release_cached_memory)."""
    return library().dabry_release_cached_memory(device)


class PinnedArray:
    """Page-locks a C-contiguous numpy array in place (dabry_host_register) for as long as this object lives, so that
    its H2D copies run at the PCIe rate.  The array itself is untouched and stays usable as plain numpy."""

    def __init__(self, array):
        if not (isinstance(array, np.ndarray) and array.flags['C_CONTIGUOUS'] and array.flags['WRITEABLE']):
            raise ValueError('pinning needs a writeable C-contiguous numpy array')
        self.array = array
        self._lib = library()
        self._ptr = array.ctypes.data
        rc = self._lib.dabry_host_register(self._ptr, array.nbytes)
        if rc != 0:
            self._ptr = None
            raise NativeError(rc, 'dabry_host_register failed (no CUDA device, or the range is already pinned)')

    def release(self):
        if getattr(self, '_ptr', None):
            self._lib.dabry_host_unregister(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Comm:
    """Owns one `dabry_comm*`: the library's own communicator over the GPUs of a job (NCCL inside libdabry_b200.so).
    `Comm.unique_id()` on rank 0, ship the bytes to the other ranks, then `Comm(device, rank, world, id)` everywhere
    (collective)."""

    @staticmethod
    def unique_id():
        lib = library()
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = lib.dabry_comm_unique_id(buf)
        if rc != 0:
            raise NativeError(rc, (lib.dabry_comm_last_error() or b'dabry_comm_unique_id failed').decode())
        return buf.raw

    def __init__(self, device, rank, world, unique_id):
        self._lib = library()
        self._c = C.c_void_p()
        self.rank, self.world, self.device = rank, world, device
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        rc = self._lib.dabry_comm_create(device, rank, world, buf, C.byref(self._c))
        if rc != 0:
            raise NativeError(rc, (self._lib.dabry_comm_last_error() or b'dabry_comm_create failed').decode())

    @property
    def pointer(self):
        return self._c

    def close(self):
        if getattr(self, '_c', None):
            self._lib.dabry_comm_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Handle:
    """Owns one `dabry_solver*`."""

    def __init__(self, config: Config, keepalive=()):
        self._lib = library()
        self._h = _H()
        self._keep = keepalive
        t0 = _now()
        rc = self._lib.dabry_create(C.byref(config), C.byref(self._h))
        _note('dabry_create', t0)
        if rc != 0:
            msg = self._lib.dabry_last_error(None)
            raise NativeError(rc, msg.decode() if msg else 'dabry_create failed')
        self.n_time = config.n_time

    def close(self):
        if getattr(self, '_h', None):
            t0 = _now()
            self._lib.dabry_destroy(self._h)
            _note('dabry_destroy', t0)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name, *args):
        t0 = _now()
        rc = getattr(self._lib, name)(self._h, *args)
        _note(name, t0)
        if rc != 0:
            msg = self._lib.dabry_last_error(self._h)
            raise NativeError(rc, msg.decode() if msg else name)

    # ---- typed helpers ----
    def num_sites(self):
        n = C.c_int64()
        self.call('dabry_num_sites', C.byref(n))
        return n.value

    def sites(self, first=0, count=None):
        count = self.num_sites() - first if count is None else count
        out = np.zeros(count, dtype=SITE_RECORD_DTYPE)
        if count:
            self.call('dabry_get_sites', out.ctypes.data_as(C.c_void_p), first, count)
        return out

    def trajs(self, first, count, want=('times', 'states', 'costates', 'controls', 'cost')):
        T = self.n_time
        shapes = {'times': (count, T), 'states': (count, T, 2), 'costates': (count, T, 2),
                  'controls': (count, T, 2), 'cost': (count, T)}
        out = {k: (np.empty(shapes[k]) if k in want else None) for k in shapes}
        if count:
            self.call('dabry_get_trajs', first, count, dptr(out['times']), dptr(out['states']),
                      dptr(out['costates']), dptr(out['controls']), dptr(out['cost']))
        return out

    def next_nb(self, first, count):
        out = np.empty((count, self.n_time), dtype=np.int32)
        if count:
            self.call('dabry_get_next_nb', first, count, iptr(out))
        return out

    def arrivals(self, first, count):
        idx = np.empty(count, dtype=np.int32)
        cost = np.empty(count)
        if count:
            self.call('dabry_get_arrivals', first, count, iptr(idx), dptr(cost))
        return idx, cost

    def solution_sites(self, bound):
        """Ascending slots of the sites with an arrival cost <= bound (device-side compaction)."""
        n = C.c_int64()
        cap = 1 << 16
        out = np.empty(cap, dtype=np.int32)
        self.call('dabry_get_solution_sites', float(bound), iptr(out), cap, C.byref(n))
        if n.value > cap:
            out = np.empty(n.value, dtype=np.int32)
            self.call('dabry_get_solution_sites', float(bound), iptr(out), n.value, C.byref(n))
        return out[:n.value].astype(np.int64)

    def solution(self):
        site = C.c_int64()
        cost = C.c_double()
        self.call('dabry_get_solution', C.byref(site), C.byref(cost))
        return site.value, cost.value

    def attach_comm(self, comm):
        self.call('dabry_attach_comm', comm.pointer if comm is not None else None)
        self._keep = tuple(self._keep) + (comm,)

    def global_solution(self):
        cost, owner, steps = C.c_double(), C.c_int32(), C.c_uint64()
        self.call('dabry_get_global_solution', C.byref(cost), C.byref(owner), C.byref(steps))
        return cost.value, owner.value, steps.value

    def stats(self):
        s = Stats()
        self.call('dabry_get_stats', C.byref(s))
        return s.as_dict()

    def flow_max_norm(self):
        v = C.c_double()
        self.call('dabry_get_flow_max_norm', C.byref(v))
        return v.value

    def cost_map(self, nx, ny, full=False):
        out = np.empty((nx, ny))
        self.call('dabry_get_cost_map', 1 if full else 0, dptr(out))
        return out

    def time_upper_bound(self, x_init, t_start, t_end):
        x = f64(np.asarray(x_init, dtype=float).reshape(2))
        out = np.empty(2)
        self.call('dabry_time_upper_bound', dptr(x), float(t_start), float(t_end), dptr(out))
        return float(out[0]), float(out[1])

    def eval_rhs(self, t, y):
        t, y = f64(np.atleast_1d(t)), f64(np.atleast_2d(y))
        dy = np.empty_like(y)
        self.call('dabry_eval_rhs', t.shape[0], dptr(t), dptr(y), dptr(dy))
        return dy

    def eval_flow(self, t, x):
        t, x = f64(np.atleast_1d(t)), f64(np.atleast_2d(x))
        w = np.empty((t.shape[0], 2))
        jac = np.empty((t.shape[0], 2, 2))
        self.call('dabry_eval_flow', t.shape[0], dptr(t), dptr(x), dptr(w), dptr(jac))
        return w, jac
