"""Parity of the pieces of the hot path in isolation, and of the two non-reference variants:

* `dabry_eval_rhs` / `dabry_eval_flow` (device A3/A4: augmented RHS, trilinear sampling, WrapperFF scalings, device
  compute_derivatives) against the numpy restatement of the reference (`pb.augsys_dyn_timeopt_*`, `ff.value/d_value`)
  at random points including points outside the grid and outside the time window;
* integrator RK4_FIXED against the oracle's scipy OdeSolver of the same stepper;
* dtype F32 (flow grid stored in fp32) against the fp64 run, tolerance 1e-4 (BASELINE.json north_star);
* edge cases: a single root, a fan that never reaches the target, capacity growth, bad arguments.

Every test body runs twice: on the GPU through libdabry_b200.so (`-m gpu`) and on tests/hostsim (CPU tier).
"""
import os

import numpy as np
import pytest

import parity_util as pu
from dump import solver_record


def _both(fn):
    """Register `fn(...)` as a hostsim test and as a gpu test."""
    def host(hostsim):
        fn()

    def gpu(cuda_lib):
        fn()
    host.__name__ = fn.__name__ + '_hostsim'
    gpu.__name__ = fn.__name__ + '_gpu'
    host.__doc__ = gpu.__doc__ = fn.__doc__
    return host, pytest.mark.gpu(gpu)


def _synthetic(gen, shape, **kw):
    from dabry_b200 import synthetic
    from dabry_b200.flowfield import DiscreteFF
    from dabry_b200.misc import Coords
    from dabry_b200.obstacle import CircleObs
    from dabry_b200.penalty import CirclePenalty
    from dabry_b200.problem import NavigationProblem
    sc = getattr(synthetic, gen)(*shape)
    ff = DiscreteFF(sc.values, sc.bounds, Coords.GCS if sc.coords == 'gcs' else Coords.CARTESIAN)
    obstacles = [CircleObs(np.array(c), float(r)) for c, r in sc.circle_obstacles]
    penalty = CirclePenalty(np.array(sc.circle_penalty[0]), *map(float, sc.circle_penalty[1:])) \
        if sc.circle_penalty else None
    pb = NavigationProblem(ff, sc.x_init, sc.x_target, sc.srf, bl=sc.bl, tr=sc.tr, obstacles=obstacles,
                           penalty=penalty).rescale()
    return pb, sc


def _random_points(pb, n, seed, margin=0.15):
    rng = np.random.default_rng(seed)
    span = pb.tr - pb.bl
    x = pb.bl - margin * span + rng.uniform(0, 1, (n, 2)) * (1 + 2 * margin) * span
    theta = rng.uniform(0, 2 * np.pi, n)
    p = rng.uniform(0.5, 1000., n)[:, None] * np.stack((np.cos(theta), np.sin(theta)), -1)
    return x, p


def rhs_and_flow_parity():
    """A3 + A4 + A5 + A6 on the device == numpy restatement, planar/spherical, with a penalty, to 1e-12."""
    from dabry_b200.solver_ef import SolverEFResampling
    for gen, shape, seed in (('planar_waves', (7, 40, 33), 1), ('era5_like', (6, 60, 31), 2),
                             ('regional_constrained', (5, 41, 23), 3)):
        pb, sc = _synthetic(gen, shape)
        solver = SolverEFResampling(pb, 1.0, n_costate_sectors=4, max_depth=1)
        x, p = _random_points(pb, 300, seed)
        t_lo, t_hi = pb.model.ff.t_start, pb.model.ff.t_end
        t = np.random.default_rng(seed).uniform(t_lo - 0.2 * (t_hi - t_lo), t_hi + 0.2 * (t_hi - t_lo), 300)
        dy = solver.dyn_augsys_device(t, np.hstack((x, p)))
        ref = np.array([solver.dyn_augsys(t[i], np.hstack((x[i], p[i]))) for i in range(300)])
        scale = np.maximum(1., np.abs(ref).max(axis=0))
        assert np.max(np.abs(dy - ref) / scale) < 1e-12, gen
        w, jac = solver._ensure_handle().eval_flow(t, x)
        w_ref = np.array([pb.model.ff.value(t[i], x[i]) for i in range(300)])
        j_ref = np.array([pb.model.ff.d_value(t[i], x[i]) for i in range(300)])
        assert np.max(np.abs(w - w_ref)) < 1e-12 * max(1., np.abs(w_ref).max())
        assert np.max(np.abs(jac - j_ref)) < 1e-12 * max(1., np.abs(j_ref).max())
        # max |w| over the nodes, computed by the repack kernel == numpy on the wrapped values (solver_ef.py:459)
        assert solver._ff_max_norm == np.max(np.linalg.norm(pb.model.ff.values, axis=-1))
        solver.close()


def analytic_rhs_parity():
    """Analytic leaves (A17 and the other catalogue fields) through the expression-tree lowering."""
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    for name in ('gyre', 'moving_vortex', 'three_vortices', 'trap', 'big_rankine', 'point_symmetric_techy2011',
                 'chertovskih', 'linear', 'four_vortices', 'three_obstacles'):
        pb = NavigationProblem.from_name(name).rescale()
        solver = SolverEFResampling(pb, 1.0, n_costate_sectors=4, max_depth=1)
        x, p = _random_points(pb, 200, 7)
        t = np.random.default_rng(5).uniform(-0.1, 1.2, 200)
        dy = solver.dyn_augsys_device(t, np.hstack((x, p)))
        ref = np.array([solver.dyn_augsys(t[i], np.hstack((x[i], p[i]))) for i in range(200)])
        scale = np.maximum(1., np.abs(ref).max(axis=0))
        assert np.max(np.abs(dy - ref) / scale) < 1e-11, name
        solver.close()


def catalogue_fields_match_reference():
    """The catalogue fields of flowfield.py:565-1134 (TwoSectors, TSEqual, LinearFFT, RadialGauss(T), BandGauss, Band,
    GyreMSEAS, and a scaled sum of them) as device leaves against golden values of the UNMODIFIED reference's classes
    (oracle/gen_field_golden.py -> tests/golden/analytic_fields.npz).  Tolerance: 1e-12 for the analytic Jacobians; the
    finite-difference ones (dx = 1e-6 as written) amplify one ulp of exp / sin / cos 1e6 times -> 1e-8 absolute on
    Jacobians of magnitude <= 25.  The host classes of this package must reproduce the reference bit for bit."""
    import dabry_b200.flowfield as mine
    from dabry_b200 import _lowering as low
    from gen_field_golden import build, field_specs
    g = np.load(os.path.join(pu.GOLDEN_DIR, 'analytic_fields.npz'))
    specs = field_specs()
    for name in specs:
        if name.startswith('_'):
            continue
        ff = build(mine, name, specs)
        t, x, w_ref, jac_ref = g[name + '_t'], g[name + '_x'], g[name + '_w'], g[name + '_jac']
        w_host = np.array([np.asarray(ff.value(t[i], x[i]), dtype=float) for i in range(len(t))])
        jac_host = np.array([np.asarray(ff.d_value(t[i], x[i]), dtype=float) for i in range(len(t))])
        assert np.array_equal(w_host, w_ref) and np.array_equal(jac_host, jac_ref), ('host mirror', name)
        handle = low.flow_evaluator(ff)
        try:
            w, jac = handle.eval_flow(t, x)
        finally:
            handle.close()
        fd = name in ('band_gauss', 'band', 'gyre_mseas', 'sum_scaled')
        assert np.max(np.abs(w - w_ref)) <= 1e-13 * max(1., np.abs(w_ref).max()), (name, np.max(np.abs(w - w_ref)))
        tol = 1e-8 if fd else 1e-12
        assert np.max(np.abs(jac - jac_ref)) <= tol * max(1., np.abs(jac_ref).max()), (name, np.max(np.abs(jac - jac_ref)))
    with pytest.raises(mine.LoweringError):  # no Jacobian in the reference either (flowfield.py:836-857)
        mine.DoubleGyreDampedFF(0.5, 0.5, 1., 1., 1., 0.5, 0.5, 0.3, 0.3).lower()
    # and through a whole solve: a Gaussian jet across the route plus the unsteady double gyre, against the oracle
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    from solver_port import Port
    ff = mine.BandGaussFF(np.array([0.5, 0.5]), np.array([0.2, 1.]), 0.6, 0.12) + 0.05 * mine.GyreMSEASFF()
    pb = NavigationProblem(ff, np.array([0.1, 0.5]), np.array([0.9, 0.5]), 1., bl=np.array([0., 0.]),
                           tr=np.array([1., 1.]), name='jet_and_gyre')
    kw = dict(max_depth=2, n_costate_sectors=16, n_time=51)
    port = Port(pb, 1.2, **kw).run_resampling()
    solver = SolverEFResampling(pb, 1.2, **kw)
    solver.solve()
    rec, ref = solver_record(solver), pu.port_record(port)
    pu.compare_sites_record(rec, ref, tol=1e-7)  # finite-difference Jacobians: 1 ulp of exp / sin -> 1e-10 per evaluation
    assert solver.native_stats()['rk_steps'] == port.n_rk
    solver.close()


def rk4_fixed_matches_oracle():
    """Integrator RK4_FIXED == the oracle's scipy OdeSolver with the same stepper (events, obstacles, sphere)."""
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    from solver_port import Port
    for name in ('gyre', 'obstacle', 'spherical_geodesic'):
        pb = NavigationProblem.from_name(name).rescale()
        kw = dict(max_depth=3, n_costate_sectors=24, n_time=101)
        port = Port(pb, integrator='rk4', **kw).run_resampling()
        # same horizon bit for bit (the solver would derive its own on the device): step counts are compared exactly
        solver = SolverEFResampling(pb, port.total_duration, integrator='rk4', **kw)
        solver.solve()
        mine, ref = solver_record(solver, full=True), pu.port_record(port)
        pu.compare_sites_record(mine, ref)
        pu.compare_full_trajectories(mine, ref)
        stats = solver.native_stats()
        assert stats['rk_steps'] == port.n_rk and stats['ivp_calls'] == port.n_ivp
        assert stats['rhs_evals'] == 4 * port.n_rk + port.n_ivp
        solver.close()


def f32_grid_within_tolerance():
    """dtype F32: same extremal counts and arrival cost, endpoints within 1e-4 of the fp64 run."""
    from cases import CASES, build_problem, solver_args
    from dabry_b200.solver_ef import SolverEFResampling
    for key in ('era5_like_R', 'planar_waves_R'):
        case = CASES[key]
        pb = build_problem(case, pu.package_api())
        args, kw = solver_args(case, pb)
        a = SolverEFResampling(*args, **kw)
        a.solve()
        b = SolverEFResampling(*args, dtype='f32', **kw)
        b.solve()
        assert list(a.sites.keys()) == list(b.sites.keys())
        assert a.solution_cost == pytest.approx(b.solution_cost, rel=1e-4)
        for name, sa in a.sites.items():
            sb = b.sites[name]
            assert len(sa.traj) == len(sb.traj) and sa.closure_reason == sb.closure_reason
            assert np.allclose(sa.traj.states[-1], sb.traj.states[-1], rtol=1e-4, atol=1e-4)
        a.close()
        b.close()


def trimming_cost_map_matches_oracle():
    """K4: the device rasteriser (bounding-box triangles + atomicMin) against the oracle's whole-grid numpy
    rasterisation (cost_map_triangle, solver_ef.py:860-889) on every sub-frame of a trimming solve, and
    `cost_map_triangle()` over all sites.  Interior cells must agree in pattern and value; the outermost line of cells is
    exempt from the pattern check: it is decided by the sign of the 1e-17 residue of extremals sliding on the frame
    (oracle/tie_sensitivity.py), where it is finite it must still carry the oracle's value."""

    def same_map(mine, ref, where):
        inner = (slice(1, -1), slice(1, -1))
        assert np.array_equal(np.isfinite(mine[inner]), np.isfinite(ref[inner])), where
        both = np.isfinite(mine) & np.isfinite(ref)
        if both.any():
            assert np.max(np.abs(mine[both] - ref[both])) < 1e-10, where
        return int(both.sum())

    from cases import CASES, build_problem, solver_args
    from dabry_b200.solver_ef import SolverEFTrimming
    from solver_port import Port
    for key in ('regional_constrained_T', 'big_rankine_T'):
        case = CASES[key]
        pb = build_problem(case, pu.package_api())
        args, kw = solver_args(case, pb)
        port = Port(*args, **kw).run_trimming()
        solver = SolverEFTrimming(*args, **kw)
        solver.setup()
        n_finite = 0
        for i in range(solver.n_subframes):  # the partial cost map of every sub-frame (solver_ef.py:825-828)
            solver.step()
            n_finite += same_map(solver._partial_cost_map, port.cost_maps[i], (key, i))
        assert n_finite > 500
        solver.check_solutions()
        assert list(solver.sites) == list(port.sites)
        same_map(solver.cost_map_triangle(), port.cost_map(port.sites.values(), 0, port.n_time - 1), (key, 'full'))
        solver.close()


def edge_cases():
    from dabry_b200 import _native as nat
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    pb = NavigationProblem.from_name('linear').rescale()
    # one root: the ring closes on itself, nothing to resample
    s = SolverEFResampling(pb, n_costate_sectors=1, max_depth=3)
    s.solve()
    assert len(s.sites) == 1 and list(s.sites)[0] == ''.rjust(0, 'A') + s.site_mngr.name_from_index(0)
    s.close()
    # a horizon too short to arrive: no solution, NaN cost, success False
    s = SolverEFResampling(pb, 0.05, max_depth=2)
    s.solve()
    assert not s.success and s.solution_site is None and np.isnan(s.solution_cost) and len(s.solution_sites) == 0
    s.close()
    # tiny initial capacity: the site store grows on demand and the result does not change
    big = SolverEFResampling(pb, max_depth=6)
    big.solve()
    small = SolverEFResampling(pb, max_depth=6, max_sites=31)
    small.solve()
    assert list(big.sites) == list(small.sites)
    assert small.native_stats()['capacity'] >= len(small.sites)
    for n in big.sites:
        assert np.array_equal(big.sites[n].traj.states, small.sites[n].traj.states)
    big.close()
    small.close()
    # argument errors surface as exceptions with the library's message
    with pytest.raises(ValueError):
        SolverEFResampling(pb, integrator='euler')
    bad = SolverEFResampling(pb, max_depth=2, root_range=(20, 40))
    with pytest.raises(nat.NativeError):
        bad.setup()


test_rhs_and_flow_parity_hostsim, test_rhs_and_flow_parity_gpu = _both(rhs_and_flow_parity)
test_analytic_rhs_parity_hostsim, test_analytic_rhs_parity_gpu = _both(analytic_rhs_parity)
test_catalogue_fields_hostsim, test_catalogue_fields_gpu = _both(catalogue_fields_match_reference)
test_rk4_fixed_matches_oracle_hostsim, test_rk4_fixed_matches_oracle_gpu = _both(rk4_fixed_matches_oracle)
test_f32_grid_within_tolerance_hostsim, test_f32_grid_within_tolerance_gpu = _both(f32_grid_within_tolerance)
test_edge_cases_hostsim, test_edge_cases_gpu = _both(edge_cases)
test_trimming_cost_map_hostsim, test_trimming_cost_map_gpu = _both(trimming_cost_map_matches_oracle)


def test_chunked_free_arcs_are_bit_identical(hostsim, monkeypatch):
    """The free-arc kernel can be left after any number of RK steps and re-entered (IvpCarry): a run in chunks of 7
    steps with re-packing of the running extremals must equal the one-launch run bit for bit."""
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    pb = NavigationProblem.from_name('moving_vortex').rescale()
    runs = []
    for chunk in ('0', '7'):
        monkeypatch.setenv('DABRY_FREE_CHUNK', chunk)
        s = SolverEFResampling(pb, max_depth=5)
        s.solve()
        runs.append((list(s.sites), {n: (site.traj.states.copy(), site.closure_reason) for n, site in s.sites.items()},
                     s.native_stats()['rk_steps'], s.native_stats()['kernel_launches']))
        s.close()
    assert runs[0][0] == runs[1][0] and runs[0][2] == runs[1][2]
    assert runs[1][3] > runs[0][3]
    for n in runs[0][0]:
        assert np.array_equal(runs[0][1][n][0], runs[1][1][n][0]) and runs[0][1][n][1] == runs[1][1][n][1]


def pinned_inputs_same_result():
    """DiscreteFF.pin_memory (dabry_host_register) changes where the flow samples live, not what is computed: a solve
    from the page-locked array equals the solve from the pageable one bit for bit, and the array stays plain numpy."""
    from dabry_b200.solver_ef import SolverEFResampling
    runs = []
    for pin in (False, True):
        pb, _ = _synthetic('era5_like', (5, 72, 37))
        inner = pb.model.ff.ff
        if pin:
            before = inner.values.copy()
            assert inner.pin_memory() is inner and inner.pin_memory() is inner  # idempotent
            assert np.array_equal(inner.values, before)
        s = SolverEFResampling(pb, 0.25, n_costate_sectors=64, max_depth=2)
        s.solve()
        runs.append((list(s.sites), np.stack([site.traj.states[-1] for site in s.sites.values()]), s.solution_cost))
        s.close()
    assert runs[0][0] == runs[1][0]
    assert np.array_equal(runs[0][1], runs[1][1])
    assert runs[0][2] == runs[1][2] or (np.isnan(runs[0][2]) and np.isnan(runs[1][2]))


test_pinned_inputs_same_result_hostsim, test_pinned_inputs_same_result_gpu = _both(pinned_inputs_same_result)


def block_cache_reuse_is_transparent():
    """Device blocks parked by a closed solver are handed to the next one (DeviceBlockCache): interleaved solvers of
    different shapes give the same results as their first run, and dabry_release_cached_memory empties the cache."""
    from dabry_b200 import _native as nat
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling

    def run(name, sectors, depth):
        pb = NavigationProblem.from_name(name).rescale()
        s = SolverEFResampling(pb, NavigationProblem.time_bound(name), n_costate_sectors=sectors, max_depth=depth,
                               max_sites=1 << 17)  # 2^17 sites: every per-sample array is above the 1 MB cache floor
        s.solve()
        out = (list(s.sites), s.solution_cost, np.stack([site.traj.states[-1] for site in s.sites.values()]))
        s.close()
        return out

    shapes = [('linear', 30, 4), ('moving_vortex', 60, 3), ('linear', 30, 4), ('three_vortices', 45, 3),
              ('moving_vortex', 60, 3)]
    first = {}
    for shape in shapes:
        res = run(*shape)
        if shape in first:
            assert res[0] == first[shape][0] and res[1] == first[shape][1]
            assert np.array_equal(res[2], first[shape][2])
        else:
            first[shape] = res
    assert nat.release_cached_memory(0) == 0
    again = run(*shapes[0])
    assert again[0] == first[shapes[0]][0] and np.array_equal(again[2], first[shapes[0]][2])


test_block_cache_reuse_is_transparent_hostsim, test_block_cache_reuse_is_transparent_gpu = \
    _both(block_cache_reuse_is_transparent)


def streamed_grid_upload_is_bit_identical(monkeypatch, exact=True):
    """A page-locked unsteady grid is uploaded in the background in four groups of frames; the first solve packs the
    record frames group by group and parks every free arc (IvpCarry) at the time the frames on the device run out.
    Results, step counts and max |w| must equal the one-shot upload bit for bit; so must a Trimming solve and the
    evaluation hooks, which wait for the whole grid."""
    from dabry_b200.solver_ef import SolverEFResampling, SolverEFTrimming
    monkeypatch.setenv('DABRY_HOSTSIM_STREAMED', '1')  # tests/hostsim: treat host arrays as page-locked
    monkeypatch.setenv('DABRY_B200_WIDE_BATCH', '0')   # one build of the arc kernels on both sides of the comparison
    for cls, kw in ((SolverEFResampling, dict(n_costate_sectors=96, max_depth=3)),
                    (SolverEFTrimming, dict(n_costate_sectors=48, max_depth=3)),
                    (SolverEFResampling, dict(n_costate_sectors=64, max_depth=2, dtype='f32')),
                    (SolverEFResampling, dict(n_costate_sectors=64, max_depth=2, integrator='rk4'))):
        runs = []
        for streamed in ('0', '2'):
            monkeypatch.setenv('DABRY_B200_STREAM_UPLOAD', streamed)
            pb, _ = _synthetic('era5_like', (12, 72, 37))
            pb.model.ff.ff.pin_memory()
            s = cls(pb, 0.25, **kw)
            s.solve()
            st = s.native_stats()
            x = np.array([[0.1, -0.9, 0.7]]).T * np.ones((1, 2)) * 0.5
            flow = s._ensure_handle().eval_flow(np.array([0.01, 0.1, 0.2]), x)
            runs.append((list(s.sites), {n: (site.traj.states.copy(), site.traj.costates.copy(), site.closure_reason)
                                        for n, site in s.sites.items()},
                         st['rk_steps'], st['ivp_calls'], s.solution_cost, s._ff_max_norm, flow, st['kernel_launches']))
            s.close()
        a, b = runs
        assert a[0] == b[0] and a[2:4] == b[2:4] and a[5] == b[5]
        assert a[4] == b[4] or (np.isnan(a[4]) and np.isnan(b[4]))
        for n in a[0]:
            if exact:
                assert np.array_equal(a[1][n][0], b[1][n][0], equal_nan=True), n
                assert np.array_equal(a[1][n][1], b[1][n][1], equal_nan=True), n
            else:  # two builds of the kernel: ptxas contracts multiply-adds differently (<= 1 ulp per operation)
                assert np.allclose(a[1][n][0], b[1][n][0], rtol=1e-9, atol=1e-9, equal_nan=True), n
            assert a[1][n][2] == b[1][n][2]
        assert np.array_equal(a[6][0], b[6][0]) and np.array_equal(a[6][1], b[6][1])
        assert b[7] > a[7]  # the streamed run really went through the windowed launches
    # the step-by-step API on a streamed grid (setup / step / check_solutions), then a second solve on the same handle
    monkeypatch.setenv('DABRY_B200_STREAM_UPLOAD', '2')
    pb, _ = _synthetic('era5_like', (12, 72, 37))
    pb.model.ff.ff.pin_memory()
    one = SolverEFResampling(pb, 0.25, n_costate_sectors=96, max_depth=3)
    one.solve()
    two = SolverEFResampling(pb, 0.25, n_costate_sectors=96, max_depth=3)
    two.setup()
    for _ in range(3):
        two.step()
    two.check_solutions()
    assert list(one.sites) == list(two.sites)
    first = {n: site.traj.states.copy() for n, site in one.sites.items()}
    for n, site in two.sites.items():
        assert np.array_equal(first[n], site.traj.states, equal_nan=True), n
    one.solve()  # the grid is resident now: one-launch free arcs
    for n, site in one.sites.items():
        if exact:
            assert np.array_equal(first[n], site.traj.states, equal_nan=True), n
        else:
            assert np.allclose(first[n], site.traj.states, rtol=1e-9, atol=1e-9, equal_nan=True), n
    one.close()
    two.close()


def test_streamed_upload_defers_starts_whose_probe_outruns_the_frames(hostsim, monkeypatch):
    """select_initial_step probes the right-hand side at t0 + h0, a distance the step limit does not bound.  With a grid
    whose frames span 2 % of the horizon the probe of every root falls beyond the frames of the first windows: the start
    is deferred (IVP_DEFERRED) to the window in which the data is there, and the results still equal the one-shot
    upload bit for bit."""
    from dabry_b200 import synthetic
    from dabry_b200.flowfield import DiscreteFF
    from dabry_b200.misc import Coords
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    monkeypatch.setenv('DABRY_HOSTSIM_STREAMED', '1')
    runs = []
    for streamed in ('0', '2'):
        monkeypatch.setenv('DABRY_B200_STREAM_UPLOAD', streamed)
        sc = synthetic.era5_like(8, 72, 37)
        bounds = sc.bounds.copy()
        bounds[0, 1] = bounds[0, 0] + 0.02 * (bounds[0, 1] - bounds[0, 0])  # eight frames inside the first 2 %
        ff = DiscreteFF(sc.values, bounds, Coords.GCS)
        ff.pin_memory()
        pb = NavigationProblem(ff, sc.x_init, sc.x_target, sc.srf, bl=sc.bl, tr=sc.tr).rescale()
        s = SolverEFResampling(pb, 0.25, n_costate_sectors=48, max_depth=2)
        s.solve()
        st = s.native_stats()
        runs.append((list(s.sites), {n: site.traj.states.copy() for n, site in s.sites.items()}, st['rk_steps'],
                     st['ivp_calls'], st['kernel_launches']))
        s.close()
    a, b = runs
    assert a[0] == b[0] and a[2:4] == b[2:4] and b[4] > a[4]
    for n in a[0]:
        assert np.array_equal(a[1][n], b[1][n], equal_nan=True), n


def test_streamed_grid_upload_is_bit_identical_hostsim(hostsim, monkeypatch):
    streamed_grid_upload_is_bit_identical(monkeypatch)


@pytest.mark.gpu
def test_streamed_grid_upload_is_bit_identical_gpu(cuda_lib, monkeypatch):
    streamed_grid_upload_is_bit_identical(monkeypatch, exact=False)


def solution_sites_compacted_on_device():
    """check_solutions keeps the sites with `not cost > 1.15 * best` (solver_ef.py:681-684): the device-side compaction
    (dabry_get_solution_sites) equals that expression on the per-site arrival arrays, which are only read back here."""
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    for name, depth in (('linear', 6), ('three_vortices', 5), ('moving_vortex', 5)):
        pb = NavigationProblem.from_name(name).rescale()
        s = SolverEFResampling(pb, NavigationProblem.time_bound(name), max_depth=depth)
        s.solve()
        assert s.success and s._arrivals is None  # nothing per-site has crossed to the host yet
        best = s.solution_cost
        idx, cost = s._arrival_index, s._arrival_cost
        expect = np.nonzero((idx >= 0) & ~(cost > 1.15 * best))[0]
        assert len(expect) >= 1 and np.array_equal(s._solution_slots, expect)
        assert {site.name for site in s.solution_sites} == {s._site_by_slot(int(k)).name for k in expect}
        assert cost[idx >= 0].min() == best
        none = s._ensure_handle().solution_sites(best * 0.5)
        assert none.shape == (0,)
        s.close()


test_solution_sites_compacted_on_device_hostsim, test_solution_sites_compacted_on_device_gpu = \
    _both(solution_sites_compacted_on_device)


def from_ff_device_sampling():
    """DiscreteFF.from_ff(device=...) (one dabry_eval_flow launch over all nodes) against the reference's node-by-node
    Python loop (flowfield.py:257-309): same bounds and shapes, values and Jacobians to 1e-12; a field that cannot be
    lowered raises instead of being sampled on the host."""
    from dabry_b200.flowfield import DiscreteFF, GyreFF, FlowField, LoweringError
    from dabry_b200.problem import NavigationProblem
    pb = NavigationProblem.from_name('moving_vortex')
    for ff, box, kw in ((pb.model.ff, (pb.bl, pb.tr), dict(nx=14, ny=11, nt=5)),
                        (GyreFF(0.5, 0.5, 2., 2., 1.), (np.zeros(2), np.ones(2)), dict(nx=12, ny=9))):
        host = DiscreteFF.from_ff(ff, box, **kw)
        dev = DiscreteFF.from_ff(ff, box, device=0, **kw)
        assert host.values.shape == dev.values.shape and np.array_equal(host.bounds, dev.bounds)
        scale = np.abs(host.values).max()
        assert np.allclose(host.values, dev.values, rtol=1e-12, atol=1e-12 * scale)
        assert np.allclose(host.grad_values, dev.grad_values, rtol=1e-12, atol=1e-12 * np.abs(host.grad_values).max())
        assert dev.coords == host.coords

    class Opaque(FlowField):
        def value(self, t, x):
            return np.zeros(2)

        def d_value(self, t, x):
            return np.zeros((2, 2))

    with pytest.raises(LoweringError):
        DiscreteFF.from_ff(Opaque(), (np.zeros(2), np.ones(2)), nx=4, ny=4, device=0)


test_from_ff_device_sampling_hostsim, test_from_ff_device_sampling_gpu = _both(from_ff_device_sampling)


class _FakeH5Node(dict):
    """Just enough of h5py.File / Group for the exporters: attrs, create_group, create_dataset."""

    def __init__(self):
        super().__init__()
        self.attrs = {}

    def create_group(self, name):
        self[name] = _FakeH5Node()
        return self[name]

    def create_dataset(self, name, data=None):
        self[name] = np.array(data)
        return self[name]


def test_legacy_rff_and_traj_exporters(hostsim, tmp_path, monkeypatch):
    """The two legacy layouts of the reference's docs (rff: coords + data/ts/grid; trajectories: one group per
    trajectory with data/ts/controls/adjoints) from a solved front.  h5py is not installed here: the npz fallback is
    checked on disk, the h5 branch against a stand-in module that records what is written."""
    import sys
    import types
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    pb = NavigationProblem.from_name('linear').rescale()
    pb.io.set_case_dir(str(tmp_path / 'case'))
    s = SolverEFResampling(pb, NavigationProblem.time_bound('linear'), max_depth=4)
    s.solve()
    path = s.save_rff(fmt='npz')
    with np.load(path) as f:
        data, ts, grid = f['data'], f['ts'], f['grid']
        assert str(f['coords']) == 'cartesian'
    assert data.shape == (s.times.shape[0], 100, 100) and grid.shape == (100, 100, 2)
    assert np.array_equal(ts, s.times)
    value = s.cost_map_triangle()
    assert np.array_equal(np.isfinite(data[0]), np.isfinite(value))
    # the front at ts[k] is the zero level set: the reached set grows with k; the cell next to the target centre is
    # reached after the target disc is entered (the solver's arrival cost) and within the horizon
    reached = (data <= 0).sum(axis=(1, 2))
    assert reached[0] <= 1 and np.all(np.diff(reached) >= 0) and reached[-1] > 1000
    ij = np.unravel_index(np.argmin(np.linalg.norm(grid - pb.x_target, axis=-1)), (100, 100))
    step = s.times[1] - s.times[0]
    assert s.solution_cost - step <= value[ij] <= s.times[-1] - s.t_init
    assert data[-1][ij] <= 0. < data[0][ij]
    assert np.allclose(grid[0, 0], pb.bl) and np.allclose(grid[-1, -1], pb.tr)

    written = {}

    class File(_FakeH5Node):
        def __init__(self, path, mode):
            super().__init__()
            written[path] = self

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    fake = types.ModuleType('h5py')
    fake.File = File
    monkeypatch.setitem(sys.modules, 'h5py', fake)
    rff_path = s.save_rff()
    assert rff_path.endswith('rff.h5')
    f = written[rff_path]
    assert f.attrs['coords'] == 'cartesian' and set(f) == {'data', 'ts', 'grid'} and np.array_equal(f['data'], data)
    traj_path = s.save_trajs_h5()
    f = written[traj_path]
    assert len(f) == len(s.sites)
    name0, site0 = next(iter(s.sites.items()))
    g = f['0']
    assert g.attrs['coords'] == 'cartesian' and g.attrs['type'] == 'pmp' and g.attrs['info'] == name0
    assert g.attrs['last_index'] == len(site0.traj) - 1 and g.attrs['interrupted'] is False
    assert np.array_equal(g['data'], site0.traj.states) and np.array_equal(g['ts'], site0.traj.times)
    assert np.array_equal(g['adjoints'], site0.traj.costates, equal_nan=True) and g['controls'].shape == (len(site0.traj),)
    s.close()


def time_upper_bound_on_device():
    """`auto_time_upper_bound` (problem.py:193-239) on the device against scipy on the host feedback
    flights, for the 11 CI cases of the reference.  These flights run without a step limit (error-controlled steps of
    rtol = 1e-3), so their arrival time is as sharp as the integrator lets rounding be: the conditioning of each flight is
    measured on scipy itself (the same flight with the control moved by <= 1 ulp per call), and the device must agree to
    1e-12 or to 100 x that spread.  The bound that matters - the smaller of the two, which fixes `total_duration` and the
    time grid - must also match the golden record of the unmodified reference to 1e-10."""
    from cases import CI_CASES, TIME_BOUND
    from dabry_b200.feedback import GSTargetFB, HTargetFB
    from dabry_b200.problem import NavigationProblem

    class Nudged:
        def __init__(self, fb, seed):
            self.fb, self.rng = fb, np.random.default_rng(seed)

        def __call__(self, t, x):
            v = np.asarray(self.fb(t, x), dtype=float)
            k = self.rng.integers(-1, 2, size=v.shape)
            return np.where(k > 0, np.nextafter(v, np.inf), np.where(k < 0, np.nextafter(v, -np.inf), v))

    arrival = NavigationProblem._arrival_time
    for name in CI_CASES:
        pb = NavigationProblem.from_name(name).rescale()
        dev = pb._feedback_times()
        laws = (GSTargetFB(pb.model.ff, pb.srf_max, pb.x_target), HTargetFB(pb.x_target, pb.model.coords))
        host = tuple(arrival(pb.apply_feedback(fb)) for fb in laws)
        for a, b, fb in zip(dev, host, laws):
            if np.isinf(b):
                assert np.isinf(a), (name, dev, host)
                continue
            spread = max(abs(arrival(pb.apply_feedback(Nudged(fb, seed))) - b) for seed in (1, 2, 3))
            assert abs(a - b) <= max(1e-12 * max(1., abs(b)), 100. * spread), (name, dev, host, spread)
        if name not in TIME_BOUND:
            g = pu.golden(name + '_R')
            total = 1.1 * min(dev)
            if total == np.inf:
                total = 2 * pb.length_reference / pb.srf_max
            assert abs(total - float(g['total_duration'])) <= 1e-10 * max(1., float(g['total_duration'])), name


test_time_upper_bound_on_device_hostsim, test_time_upper_bound_on_device_gpu = _both(time_upper_bound_on_device)


def exporters_against_oracle():
    """SURVEY 8(f) rank 3 against the ORACLE, not against itself: the legacy rff product (`save_rff`, data[k] = V - ts[k]
    with V the device cost map of the whole front) equals what the oracle's restatement of cost_map_triangle
    (solver_ef.py:860-889) gives for its own front, cell by cell in the interior of the map (border cells depend on
    extremals sliding on the frame), and the one-file front bundle (`save_front_bundle`) holds the oracle's
    trajectories."""
    import tempfile
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    from solver_port import Port
    name, depth = 'obstacle', 5
    pb = NavigationProblem.from_name(name).rescale()
    tmp = tempfile.mkdtemp(prefix='dabry_export_')
    pb.io.set_case_dir(os.path.join(tmp, 'case'))
    s = SolverEFResampling(pb, max_depth=depth)
    s.solve()
    port = Port(pb, max_depth=depth)
    port.run_resampling()
    ref_map = port.cost_map(port.sites.values(), 0, port.n_time - 1)
    with np.load(s.save_rff(fmt='npz')) as f:
        data, ts, grid = f['data'], f['ts'], f['grid']
    assert np.array_equal(ts, port.times)
    value = data[0] + ts[0] - s.t_init  # data[k] = V + t_init - ts[k]
    inner = (slice(1, -1), slice(1, -1))
    assert np.array_equal(np.isfinite(value[inner]), np.isfinite(ref_map[inner]))
    fin = np.isfinite(ref_map[inner])
    assert fin.sum() > 1000
    assert np.max(np.abs(value[inner][fin] - ref_map[inner][fin])) <= 1e-9
    for k in (0, 37, 99):
        assert np.array_equal(data[k], data[0] + ts[0] - ts[k])
    gx = np.linspace(pb.bl[0], pb.tr[0], 100)
    assert np.array_equal(grid[:, 0, 0], gx)
    # the bundle: padded trajectories + records of every site, in insertion order
    s.save_front_bundle('front')
    with np.load(os.path.join(pb.io.trajs_dir, 'front.npz')) as f:
        names, states, records, times_grid = list(f['names']), f['states'], f['records'], f['times_grid']
    assert names == [x.name for x in port.sites.values()] and np.array_equal(times_grid, port.times)
    for i, site in enumerate(port.sites.values()):
        k0, n = site.index_t_init, len(site.t)
        assert records['index_t_init'][i] == k0 and records['n_samples'][i] == n
        assert np.max(np.abs(states[i, k0:k0 + n] - np.array(site.x))) <= 1e-9
    s.close()


test_exporters_against_oracle_hostsim, test_exporters_against_oracle_gpu = _both(exporters_against_oracle)


def test_lazy_host_jacobian_and_npz_ingest(hostsim, tmp_path):
    """DiscreteFF: the default Jacobian (flowfield.py:474-504) is computed on the host only when read - same values as
    the reference's eager constructor -, a solve never materialises it (the device differentiates at upload), a
    caller-supplied Jacobian is uploaded; `from_npz` round trip of docs/flow_format.md."""
    from dabry_b200 import synthetic
    from dabry_b200.flowfield import DiscreteFF, save_ff
    from dabry_b200.misc import Coords
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    sc = synthetic.planar_waves(5, 24, 20)
    ff = DiscreteFF(sc.values, sc.bounds, Coords.CARTESIAN)
    assert ff._grad_values is None and ff.grad_is_device_default() and not ff.has_host_jacobian()
    eager = DiscreteFF(sc.values, sc.bounds, Coords.CARTESIAN, force_no_diff=True)
    eager.compute_derivatives()
    pb = NavigationProblem(ff, sc.x_init, sc.x_target, sc.srf, bl=sc.bl, tr=sc.tr, name='lazy').rescale()
    solver = SolverEFResampling(pb, sc.total_duration, n_costate_sectors=12, max_depth=1)
    solver.solve()
    cost = solver.solution_cost
    solver.close()
    assert ff._grad_values is None, 'the solve materialised the host Jacobian'
    assert np.array_equal(ff.grad_values, eager.grad_values) and ff.grad_is_device_default()
    # the same field with the Jacobian handed over by the caller: uploaded, same result
    custom = DiscreteFF(sc.values, sc.bounds, Coords.CARTESIAN, grad_values=eager.grad_values.copy())
    assert custom.has_host_jacobian()
    pb2 = NavigationProblem(custom, sc.x_init, sc.x_target, sc.srf, bl=sc.bl, tr=sc.tr, name='custom').rescale()
    solver = SolverEFResampling(pb2, sc.total_duration, n_costate_sectors=12, max_depth=1)
    solver.solve()
    assert solver.solution_cost == cost
    solver.close()
    path = str(tmp_path / 'ff.npz')
    save_ff(ff, path)
    back = DiscreteFF.from_npz(path)
    assert np.array_equal(back.values, ff.values) and np.array_equal(back.bounds, ff.bounds) and back.coords == ff.coords
    assert back._grad_values is None and back.values.flags['C_CONTIGUOUS'] and back.values.dtype == np.float64


def test_root_fan_matches_numpy_and_shards():
    """The threaded, cached root fan equals the reference's expression 1000 (cos, sin)(linspace) bit for bit
    (solver_ef.py:465-467), and a block-cyclic shard is the matching subset of the ring."""
    from dabry_b200.solver_ef import _root_fan
    n = 1 << 17
    theta = np.linspace(0, 2 * np.pi, n, endpoint=False)
    ref = 1000 * np.stack((np.cos(theta), np.sin(theta)), -1)
    full = _root_fan(n, None, None)
    assert np.array_equal(full, ref)
    assert _root_fan(n, None, None) is full  # cached
    rank, world, block = 1, 4, 64
    part = _root_fan(n, (rank * block, n // world), (rank, world, block))
    j = np.arange(n // world)
    assert np.array_equal(part, ref[rank * block + (j // block) * (world * block) + j % block])
    small = _root_fan(30, None, None)
    th = np.linspace(0, 2 * np.pi, 30, endpoint=False)
    assert np.array_equal(small, 1000 * np.stack((np.cos(th), np.sin(th)), -1))


def test_library_exports_every_declared_symbol():
    """Every function include/dabry_b200.h declares is exported by the built CUDA library and bound by ctypes (no
    compute call: this runs without a GPU)."""
    import ctypes
    import os
    import re
    from dabry_b200 import _native as nat
    header = open(os.path.join(pu.ROOT, 'include', 'dabry_b200.h')).read()
    declared = set(re.findall(r'\b(dabry_[a-z0-9_]+)\s*\(', header))
    declared -= {'dabry_solver'}
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    assert os.path.exists(nat.DEFAULT_LIBRARY), 'build the extension: python -c "import __graft_entry__ as g; g.build()"'
    lib = ctypes.CDLL(nat.DEFAULT_LIBRARY)
    for name in declared:
        assert hasattr(lib, name), name
    lib.dabry_backend_name.restype = ctypes.c_char_p
    assert lib.dabry_backend_name() == b'cuda-sm100a'
    assert lib.dabry_abi_version() == nat.ABI_VERSION
    # ctypes mirrors of the ABI structs have the C sizes (checked against a compiled probe)
    import subprocess
    import tempfile
    probe = ('#include "dabry_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu",'
             'sizeof(dabry_config),sizeof(dabry_wrapper),sizeof(dabry_leaf),sizeof(dabry_obstacle),'
             'sizeof(dabry_penalty),sizeof(dabry_site_record),sizeof(dabry_stats));}')
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, 'probe.c')
        open(src, 'w').write(probe)
        subprocess.run(['gcc', '-I', os.path.join(pu.ROOT, 'include'), src, '-o', os.path.join(d, 'probe')], check=True)
        sizes = list(map(int, subprocess.run([os.path.join(d, 'probe')], capture_output=True, text=True).stdout.split()))
    assert sizes == [ctypes.sizeof(nat.Config), ctypes.sizeof(nat.Wrapper), ctypes.sizeof(nat.Leaf),
                     ctypes.sizeof(nat.Obstacle), ctypes.sizeof(nat.Penalty), nat.SITE_RECORD_DTYPE.itemsize,
                     ctypes.sizeof(nat.Stats)]


def test_no_cpu_fallback_without_device():
    """On a machine without a CUDA device the product library refuses to create a solver (no CPU path)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip('a GPU is present')
    except ImportError:
        pass
    from dabry_b200 import _native as nat
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    nat.use_library(None)
    pb = NavigationProblem.from_name('linear').rescale()
    solver = SolverEFResampling(pb, 1.0, max_depth=1)
    with pytest.raises(nat.NativeError) as err:
        solver.solve()
    assert err.value.code == -2  # DABRY_ERR_NO_DEVICE
