"""Test infrastructure (NOT product code): import the unmodified reference (dabry v1.1) from
/root/reference in THIS container only, under the three shims SURVEY.md §8c lists.

* stub modules ``h5py`` and ``pygrib`` (imported at module top by flowfield.py:9-11, io_manager.py:9-11,
  penalty.py:3 but used only by file loaders);
* ``numpy.infty`` (removed in NumPy 2; used at solver_ef.py:306, problem.py:234,239);
* ``sys.path`` entry for the read-only mount.

/root/reference does not exist on the GPU box: only ``oracle/gen_golden.py`` and the container-only
cross-checks in ``tests/`` (skipped when the mount is absent) may call :func:`load_reference`.
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get('DABRY_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'dabry'))


def load_reference():
    """Returns the reference's top-level package ``dabry`` (imported once)."""
    if not reference_available():
        raise RuntimeError(f'reference not mounted at {REFERENCE_ROOT}')
    import numpy
    if not hasattr(numpy, 'infty'):
        numpy.infty = numpy.inf
    for missing in ('h5py', 'pygrib'):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except ImportError:
                sys.modules[missing] = types.ModuleType(missing)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    warnings.filterwarnings('ignore')
    import dabry  # noqa
    import dabry.solver_ef  # noqa
    import dabry.problem  # noqa
    return dabry
