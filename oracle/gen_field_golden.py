"""Test infrastructure (NOT product code): golden values of the analytic catalogue fields, from the UNMODIFIED reference.

    python oracle/gen_field_golden.py          (container only: needs /root/reference)

Evaluates `value(t, x)` and `d_value(t, x)` of every catalogue field of dabry/flowfield.py:565-1134 that the device
implements as a leaf, through the reference's own classes, at seeded points -> tests/golden/analytic_fields.npz.
`field_specs()` is shared with the tests (tests/test_variants_and_hooks.py builds the same fields from this package's
classes).  LinearFFT.value has no return statement in the reference (flowfield.py:776-777): its value is recorded as
the documented expression evaluated with numpy, its Jacobian through the reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, 'tests', 'golden', 'analytic_fields.npz')
N_POINTS = 64


def field_specs():
    """name -> (class name, args, kwargs, (x lo, x hi), (t lo, t hi))"""
    nt = 5
    rng = np.random.default_rng(11)
    return {
        'two_sectors': ('TwoSectorsFF', (0.3, -0.2, 0.45), {}, (0., 1.), (0., 1.)),
        'ts_equal': ('TSEqualFF', (0.1, 0.25, 1.0), {}, (0., 1.), (0., 1.)),
        'linear_t': ('LinearFFT', (np.array([[0.1, -0.3], [0.2, 0.05]]), np.array([[0., 0.4], [-0.4, 0.]]),
                                   np.array([0.2, 0.1]), np.array([0.05, -0.02])), {}, (-1., 1.), (0., 2.)),
        'radial_gauss': ('RadialGaussFF', (np.array([0.4, 0.6]), 0.3, 0.5, 0.8), {}, (0., 1.), (0., 1.)),
        'radial_gauss_t': ('RadialGaussFFT', (np.column_stack((np.linspace(0.2, 0.7, nt), np.linspace(0.6, 0.3, nt))),
                                              np.linspace(0.2, 0.4, nt), np.linspace(0.4, 0.6, nt),
                                              np.linspace(0.5, 1.0, nt), 1.5), dict(t_start=0.25), (0., 1.), (0., 2.)),
        'band_gauss': ('BandGaussFF', (np.array([0.5, 0.5]), np.array([1., 0.5]), 0.7, 0.15), {}, (0., 1.), (0., 1.)),
        'band': ('BandFF', (np.array([0.5, 0.4]), np.array([0.3, 1.]), np.array([0.2, 0.6]), 0.4), {}, (0., 1.), (0., 1.)),
        'gyre_mseas': ('GyreMSEASFF', (), {}, (0., 2.), (0., 2.)),
        'sum_scaled': ('expr', None, {}, (0., 1.), (0., 1.)),  # 0.5 * band_gauss + two_sectors - 2. * gyre_mseas
        '_seed': int(rng.integers(1 << 30)),
    }


def build(mod, name, specs):
    if name == 'sum_scaled':
        return 0.5 * build(mod, 'band_gauss', specs) + build(mod, 'two_sectors', specs) - 2. * build(mod, 'gyre_mseas', specs)
    cls, args, kw = specs[name][:3]
    return getattr(mod, cls)(*args, **kw)


def points(name, specs):
    (xlo, xhi), (tlo, thi) = specs[name][3], specs[name][4]
    rng = np.random.default_rng(abs(hash(name)) % (1 << 30) if False else sum(map(ord, name)))
    x = rng.uniform(xlo, xhi, (N_POINTS, 2))
    t = rng.uniform(tlo - 0.1, thi + 0.1, N_POINTS)
    return t, x


def main():
    sys.path.insert(0, HERE)
    import refshim
    refshim.load_reference()
    import dabry.flowfield as ref
    specs = field_specs()
    out = {}
    for name in specs:
        if name.startswith('_'):
            continue
        ff = build(ref, name, specs)
        t, x = points(name, specs)
        if name == 'linear_t':  # the reference's value() returns None: the documented expression
            w = np.array([(ff.gradient_diff_time * t[i] + ff.gradient_init) @ (x[i] - ff.origin) + ff.value_origin
                          for i in range(N_POINTS)])
        else:
            w = np.array([np.asarray(ff.value(t[i], x[i]), dtype=float) for i in range(N_POINTS)])
        jac = np.array([np.asarray(ff.d_value(t[i], x[i]), dtype=float) for i in range(N_POINTS)])
        out[name + '_t'], out[name + '_x'], out[name + '_w'], out[name + '_jac'] = t, x, w, jac
        print('%-16s |w| max %.3g  |J| max %.3g' % (name, np.abs(w).max(), np.abs(jac).max()))
    np.savez_compressed(OUT, **out)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
