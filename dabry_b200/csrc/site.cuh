// The extremal field as the device holds it, and the per-site procedures of the reference's solver
// (dabry/solver_ef.py), each written for ONE site so that a kernel runs one thread per site.
//
//   SiteStore                 <- class Site (solver_ef.py:56-164): struct of arrays, slot = time-grid index
//   site_integrate            <- SolverEFResampling.integrate_site_to_target_index (:479-580)
//   site_connect              <- connect_to_parents (:690-699)
//   site_resample             <- compute_new_sites (:630-665), one site of the loop
//   site_spawn                <- SiteManager.site_from_parents (:261-278)
//   site_trim_distance        <- trim_distance (:594-600)
//   site_arrival              <- check_solutions (:667-685), per-site part
//   site_rasterize            <- cost_map_triangle (:860-889) + triangle_mask_and_cost (misc.py:115-139)
//   site_trim                 <- SolverEFTrimming.trim (:819-841), per-site part
#pragma once
#include "ivp.cuh"
#include "model.cuh"

// 1: the throughput build of the free-arc kernel of GRIDDED fields runs the rolled step (rk45.cuh rk_step_rolled) with
// the right-hand side inlined once and the Runge-Kutta stage values in shared memory.  0: the unrolled step with a called
// right-hand side everywhere (round 1's kernel).  The latency builds (small batches, one extremal per warp) and analytic
// fields always keep the unrolled step: they are bound by one thread's dependent chain, which the stage loop and the
// shared-memory round trips lengthen (movor 28.7 -> 30.0 ms, Gyre 2.74 -> 2.87 ms with the rolled step).
#ifndef DABRY_FAST_STEP
#define DABRY_FAST_STEP 1
#endif
// 1: stage values of the unrolled step in shared memory too (measured equal in round 1: 25.9 vs 25.5 ms)
#ifndef DABRY_K_IN_SHARED
#define DABRY_K_IN_SHARED 0
#endif
// 1: a lane whose step attempt is rejected repeats it in the warp's next round instead of inside scipy's inner loop
// (rk45.cuh rk_step DEFER) - free arcs of the one-launch builds only
#ifndef DABRY_DEFER_RETRY
#define DABRY_DEFER_RETRY 0
#endif
#ifndef DABRY_FREE_ARC_THREADS
#define DABRY_FREE_ARC_THREADS 128  // block size of the free-arc kernel (FreeArcPhase::kThreads)
#endif

namespace dabry {

struct Sample {  // 32-byte record: one aligned sector per grid sample
    double x0, x1, p0, p1;
};
struct Control {
    double u0, u1;
};

enum SiteFlags : int32_t { SITE_GHOST = 1, SITE_CONNECTED = 2, SITE_PENDING_ENTRY = 4, SITE_RUNNING = 8,
                           SITE_DEFERRED = 16 };  // DEFERRED: the arc could not start in this window (streamed grid)

struct SiteStore {
    int32_t n_time;
    int32_t capacity;
    double t_init;        // solver.t_init: cost = t - t_init (solver_ef.py:501)
    const double* times;  // [n_time] solver.times (solver_ef.py:313)
    // trajectories, [n_time][capacity] (time-major), row k holds the samples at grid index k
    Sample* xp;      // states + costates (costates NaN on constrained arcs)
    Control* ctrl;   // controls (NaN on free arcs)
    double* tt;      // sample times (res.t)
    // per site
    double* cost0;        // cost of the first sample (0 for roots, mean of the parents' for children)
    double* t_enter_obs;
    int32_t* index_t_init;
    int32_t* n_samples;   // len(site.traj)
    int32_t* closure;
    int32_t* neuter;
    int32_t* index_neutered;
    int32_t* obstacle;    // event index of the capturing obstacle (1-based into the event table), -1 = free
    int32_t* index_t_obs;
    int32_t* trigo;
    int32_t* check_next;  // index_t_check_next
    int32_t* depth;
    int32_t* flags;
    int32_t* parent_prev;
    int32_t* parent_next;
    int32_t* created_subframe;  // trimming: sub-frame during which the site was created (roots: 0)
    int32_t* valid_first;       // trimming: site belongs to sites_valid[k] for valid_first <= k <= valid_last
    int32_t* valid_last;
    int64_t* dyadic;            // SiteManager index (solver_ef.py:197-202)
    int32_t* next_nb;           // [n_time][capacity], -1 = None
    int32_t* n_target_events;   // events['target'] (non terminal): count and the first few roots
    double* t_target_events;    // [capacity][DABRY_MAX_TARGET_EVENTS]
    int32_t* ivp_status;        // status of the last solve_ivp call of the site (diagnostics)
    int32_t* arc_from;          // grid index the current integrate call started from (-1: nothing to do)
    int32_t* arr_index;         // check_solutions, kept up to date by whoever writes a sample: grid index of the cheapest
    double* arr_cost;           //   own sample strictly inside the target disc (-1 / +inf: none), and its cost
    Sample* enter;              // augmented state at the obstacle entry root, between the two integrate kernels
    // state of a free arc parked between two chunks of RK steps (IvpCarry)
    double* carry_t;
    double* carry_h;
    Sample* carry_y;
    Sample* carry_f;
    double* carry_g;            // [capacity][DABRY_MAX_EVENTS]
    int32_t* carry_eval_i;
    // time-major addressing: consecutive sites are consecutive in memory at a given grid index, so that a warp
    // of extremals advancing in lockstep writes (and the resampling scan reads) fully coalesced rows
    DABRY_HD int64_t at(int s, int k) const { return (int64_t) k * capacity + s; }
};

struct SolverParams {
    int32_t max_depth;
    int32_t mode_origin;
    int32_t follow_obstacles;  // 0: SolverEFSimple semantics - stop at the first obstacle event
    int32_t n_roots;
    int32_t cm_nx, cm_ny;  // trimming cost map (solver_ef.py:820)
    double max_dist_sq;
    double ff_max_norm;    // max |w| over the grid nodes, <= 0 disables trim_distance (solver_ef.py:459-460)
    double bl[2], tr[2];   // problem box
    double cm_inv[2];      // (cm_n - 1) / (tr - bl): reciprocal cell size of the cost map, for the bounding-box index ranges
};

DABRY_HD int site_index_t(const SiteStore& S, int s) { return S.index_t_init[s] + S.n_samples[s] - 1; }
DABRY_HD bool site_in_obs_at(const SiteStore& S, int s, int k) {
    return S.obstacle[s] >= 0 && k >= S.index_t_obs[s];
}
DABRY_HD double site_cost_at(const SiteStore& S, int s, int k) {
    return k == S.index_t_init[s] ? S.cost0[s] : S.tt[S.at(s, k)] - S.t_init;
}

DABRY_HD void site_reset(const SiteStore& S, int s) {
    S.cost0[s] = 0.;
    S.t_enter_obs[s] = quiet_nan();
    S.index_t_init[s] = 0;
    S.n_samples[s] = 1;
    S.closure[s] = CLOSURE_NONE;
    S.neuter[s] = NEUTER_NONE;
    S.index_neutered[s] = -1;
    S.obstacle[s] = -1;
    S.index_t_obs[s] = 0;
    S.trigo[s] = 1;
    S.check_next[s] = 0;
    S.depth[s] = 0;
    S.flags[s] = 0;
    S.parent_prev[s] = -1;
    S.parent_next[s] = -1;
    S.created_subframe[s] = 0;
    S.valid_first[s] = 0x7fffffff;
    S.valid_last[s] = -1;
    S.dyadic[s] = 0;
    S.n_target_events[s] = 0;
    S.ivp_status[s] = 0;
    S.arc_from[s] = -1;
    S.arr_index[s] = -1;
    S.arr_cost[s] = INFINITY;
    // next_nb rows are filled with -1 when the storage is allocated; slots are never recycled
}

// ---------------------------------------------------------------------------------------------------------
// integrate_site_to_target_index

template <int FK, bool INLINED = false>
struct AugRhs {
    const ProblemDev& P;
    DABRY_HD void operator()(double t, const double* y, double* dy) const { rhs_aug<FK, INLINED>(P, t, y, dy); }
};

template <int FK>
struct ConstrRhs {
    const ProblemDev& P;
    int obstacle;  // 0-based index into P.obstacles
    bool trigo;
    DABRY_HD void operator()(double t, const double* x, double* dx) const {
        rhs_constr<FK>(P, obstacle, trigo, t, x, dx, nullptr);
    }
};

struct FreeEvents {
    static constexpr bool kBothDirections = false;  // misc.py:34-43: every event of the solver has direction -1
    const ProblemDev& P;
    const SiteStore& S;
    int site;
    DABRY_HD int count() const { return 1 + P.n_obstacles; }
    DABRY_HD bool terminal(int e) const { return e > 0; }
    DABRY_HD double value(int e, double, const double* y) const { return event_value(P, e, y[0], y[1]); }
    DABRY_HD void record(int e, double t) {
        if (e == 0) {
            const int n = S.n_target_events[site];
            if (n < DABRY_MAX_TARGET_EVENTS) S.t_target_events[(int64_t) site * DABRY_MAX_TARGET_EVENTS + n] = t;
            S.n_target_events[site] = n + 1;
        }
    }
};

// check_solutions (solver_ef.py:667-685), the per-sample half: a sample strictly inside the target disc is a candidate
// arrival, the cheapest one is kept.  Called by every writer of a sample (roots, children, both kinds of arcs) right
// after the sample and its time are stored, so that no pass has to read the trajectories again.
DABRY_HD void arrival_note(const ProblemDev& P, const SiteStore& S, int s, int k, double x0, double x1) {
    const double d0 = x0 - P.x_target[0], d1 = x1 - P.x_target[1];
    if (add_rn(mul_rn(d0, d0), mul_rn(d1, d1)) < P.target_radius_sq) {
        const double c = site_cost_at(S, s, k);
        if (S.arr_index[s] < 0 || c < S.arr_cost[s]) {
            S.arr_index[s] = k;
            S.arr_cost[s] = c;
        }
    }
}

struct NoEvents {
    static constexpr bool kBothDirections = false;
    DABRY_HD int count() const { return 0; }
    DABRY_HD bool terminal(int) const { return false; }
    DABRY_HD double value(int, double, const double*) const { return 0.; }
    DABRY_HD void record(int, double) {}
};

template <int FK>
struct QuitEvent {
    static constexpr bool kBothDirections = false;
    const ProblemDev& P;
    int obstacle;
    bool fired;
    DABRY_HD int count() const { return 1; }
    DABRY_HD bool terminal(int) const { return true; }
    DABRY_HD double value(int, double t, const double* x) const {
        return event_quit_obstacle<FK>(P, obstacle, t, x[0], x[1]);
    }
    DABRY_HD void record(int, double) { fired = true; }
};

struct FreeSink {
    const ProblemDev& P;
    const SiteStore& S;
    int site, slot0;  // grid index of linspace element 0
    DABRY_HD void emit(int j, double t, const double* y) {
        const int64_t o = S.at(site, slot0 + j);
        S.xp[o] = Sample{y[0], y[1], y[2], y[3]};
        S.tt[o] = t;  // the control stays NaN (free arc, solver_ef.py:500): the store is pre-filled, no write here
        arrival_note(P, S, site, slot0 + j, y[0], y[1]);
    }
};

template <int FK>
struct ConstrSink {
    const ProblemDev& P;
    const SiteStore& S;
    int site, slot0, obstacle;
    bool trigo;
    DABRY_HD void emit(int j, double t, const double* x) {
        const int64_t o = S.at(site, slot0 + j);
        double dx[2], u[2];
        rhs_constr<FK>(P, obstacle, trigo, t, x, dx, u);  // dyn_constr - ff.value (solver_ef.py:536-538)
        S.xp[o] = Sample{x[0], x[1], quiet_nan(), quiet_nan()};
        S.ctrl[o] = Control{u[0], u[1]};
        S.tt[o] = t;
        arrival_note(P, S, site, slot0 + j, x[0], x[1]);
    }
};

// integrate_site_to_target_index is split in two device procedures so that the hot kernel (free arcs: every
// extremal, ~all of the RK work) stays small enough for the instruction cache and the rare, branchy obstacle
// handling runs in a second kernel over the captured sites only.
//
// Part A - the free arc (solver_ef.py:482-504): solve_ivp on the augmented system with the target / obstacle events.
// A terminal obstacle event leaves the entry record (obstacle, t_enter_obs, index_t_obs, state at entry) for part B.
template <int FK, int METHOD, bool CHUNKED, bool FAST = false>
DABRY_HD void site_free_arc(const ProblemDev& P, const SiteStore& S, int s, int index_t, unsigned& n_attempts,
                            unsigned& n_calls, int step_budget, double t_park = INFINITY, bool resume_only = false) {
    IvpCarry<4> carry;
    int idx0;
    if (CHUNKED && resume_only && !(S.flags[s] & (SITE_RUNNING | SITE_DEFERRED))) return;  // finished in an earlier window
    if (CHUNKED) S.flags[s] &= ~SITE_DEFERRED;
    if (CHUNKED && (S.flags[s] & SITE_RUNNING)) {  // re-entry after a chunk of steps
        idx0 = S.arc_from[s];
        carry.started = 1;
        carry.eval_i = S.carry_eval_i[s];
        carry.t = S.carry_t[s];
        carry.h_abs = S.carry_h[s];
        const Sample cy = S.carry_y[s], cf = S.carry_f[s];
        carry.y[0] = cy.x0; carry.y[1] = cy.x1; carry.y[2] = cy.p0; carry.y[3] = cy.p1;
        carry.f[0] = cf.x0; carry.f[1] = cf.x1; carry.f[2] = cf.p0; carry.f[3] = cf.p1;
        for (int e = 0; e <= P.n_obstacles; ++e) carry.g[e] = S.carry_g[(int64_t) s * DABRY_MAX_EVENTS + e];
    } else {
        idx0 = site_index_t(S, s);
        S.arc_from[s] = -1;
        if (idx0 >= index_t) return;
        if (index_t - idx0 + 1 <= 1) return;
        S.arc_from[s] = idx0;
        if (site_in_obs_at(S, s, idx0)) return;
        carry.started = 0;
        ++n_calls;
        // Trajectory.__add__ (trajectory.py:136-137): the events of a later arc replace those of the arcs before it, so
        // events['target'] of a site integrated sub-frame by sub-frame (Trimming) holds the roots of its LAST free arc
        S.n_target_events[s] = 0;
    }
    const int n = index_t - idx0 + 1;
    const Sample y0s = S.xp[S.at(s, idx0)];
    const double y0[4] = {y0s.x0, y0s.x1, y0s.p0, y0s.p1};
    const EvalGrid grid = make_eval_grid(S.times[idx0], S.times[index_t], n, 1);
    // FAST: rolled step, right-hand side inlined once (Dormand-Prince only: RK4 has no stage table here)
    constexpr bool kRolled = FAST && METHOD == METHOD_DOPRI5;
    AugRhs<FK, kRolled> rhs{P};
    FreeEvents ev{P, S, s};
    FreeSink sink{P, S, s, idx0};
    IvpResult<4> res;
#if defined(__CUDA_ARCH__)
    if constexpr (FAST || DABRY_K_IN_SHARED != 0) {
        // the stage values of the step live in shared memory (rk45.cuh KShared): 32 KB per block of 128 threads
        __shared__ __align__(16) double k_tile[8 * 4 * DABRY_FREE_ARC_THREADS];
        solve_ivp<4, METHOD, KShared<4, DABRY_FREE_ARC_THREADS>, false, kRolled>(
            rhs, ev, sink, grid.at(0), grid.at(n - 1), y0, grid, P.max_step, P.rtol, P.atol, res, n_attempts,
            CHUNKED ? &carry : nullptr, CHUNKED ? step_budget : 0,
            KShared<4, DABRY_FREE_ARC_THREADS>{k_tile + 2 * threadIdx.x}, t_park);
    } else
#endif
    {
        solve_ivp<4, METHOD, KRegisters<4>, (!CHUNKED && DABRY_DEFER_RETRY != 0), kRolled>(
            rhs, ev, sink, grid.at(0), grid.at(n - 1), y0, grid, P.max_step, P.rtol, P.atol, res, n_attempts,
            CHUNKED ? &carry : nullptr, CHUNKED ? step_budget : 0, KRegisters<4>(), t_park);
    }
    if (CHUNKED && res.status == IVP_DEFERRED) {  // nothing happened yet: a later window starts this arc afresh
        S.flags[s] |= SITE_DEFERRED;
        --n_calls;
        return;
    }
    S.n_samples[s] += res.n_emitted;
    if (CHUNKED && res.status == IVP_RUNNING) {
        S.flags[s] |= SITE_RUNNING;
        S.carry_eval_i[s] = carry.eval_i;
        S.carry_t[s] = carry.t;
        S.carry_h[s] = carry.h_abs;
        S.carry_y[s] = Sample{carry.y[0], carry.y[1], carry.y[2], carry.y[3]};
        S.carry_f[s] = Sample{carry.f[0], carry.f[1], carry.f[2], carry.f[3]};
        for (int e = 0; e <= P.n_obstacles; ++e) S.carry_g[(int64_t) s * DABRY_MAX_EVENTS + e] = carry.g[e];
        return;
    }
    if (CHUNKED) S.flags[s] &= ~SITE_RUNNING;
    S.ivp_status[s] = res.status;
    if (res.status == IVP_TERMINATED) {
        S.t_enter_obs[s] = res.t_term;
        S.obstacle[s] = res.term_event;
        S.index_t_obs[s] = site_index_t(S, s) + 1;  // site.index_t + len(traj_free) + 1 (:516)
        S.enter[s] = Sample{res.y_term[0], res.y_term[1], res.y_term[2], res.y_term[3]};
        S.flags[s] |= SITE_PENDING_ENTRY;
    }
}

// Part B - obstacle entry and boundary following (solver_ef.py:506-576) for a site captured during part A of this
// call, or captured earlier and not yet at index_t.
template <int FK, int METHOD>
DABRY_HD void site_obstacle_arcs(const ProblemDev& P, const SiteStore& S, int s, int index_t, unsigned& n_attempts,
                                 unsigned& n_calls) {
    const int idx0 = S.arc_from[s];
    if (idx0 < 0 || S.obstacle[s] < 0) return;
    const int n = index_t - idx0 + 1;
    const double t_start = S.times[idx0], t_end = S.times[index_t];
    const int obs = S.obstacle[s] - 1;  // index into P.obstacles
    if (S.flags[s] & SITE_PENDING_ENTRY) {
        S.flags[s] &= ~SITE_PENDING_ENTRY;
        const EvalGrid grid = make_eval_grid(t_start, t_end, n, 1);
        const int n_free = S.index_t_obs[s] - 1 - idx0;
        const double t_enter = S.t_enter_obs[s];
        const Sample ye = S.enter[s];
        double g0, g1;
        obstacle_grad(P.obstacles[obs], ye.x0, ye.x1, g0, g1);
        const double pn = sqrt(ye.p0 * ye.p0 + ye.p1 * ye.p1);
        double v0, v1;
        dyn_value<FK>(P, t_enter, ye.x0, ye.x1, -ye.p0 / pn, -ye.p1 / pn, v0, v1);
        const bool trigo = (g0 * v1 - g1 * v0) >= 0.;  // np.cross(d_value, dyn.value) (solver_ef.py:519-523)
        S.trigo[s] = trigo ? 1 : 0;
        const double sign = trigo ? 1. : -1.;
        const double d0 = -sign * g1, d1 = sign * g0;
        Flow f;
        field_eval<FK, false>(P.field, t_enter, ye.x0, ye.x1, f);
        bool tracked = false;
        if (possible_direction(f.w0, f.w1, d0, d1, P.srf)) {
            // t_eval[t_eval > t_enter_obs][0]  (solver_ef.py:528-530)
            int jn = n_free + 1;
            while (jn < n && !(grid.at(jn) > t_enter)) ++jn;
            if (jn < n) {
                const double t_next = grid.at(jn);
                EvalGrid one;
                one.t_start = t_next;
                one.t_end = t_next;
                one.step = 0.;
                one.n = 1;
                one.first = 0;
                ConstrRhs<FK> crhs{P, obs, trigo};
                NoEvents nev;
                ConstrSink<FK> csink{P, S, s, idx0 + n_free + 1, obs, trigo};
                IvpResult<2> r2;
                const double x0c[2] = {ye.x0, ye.x1};
                solve_ivp<2, METHOD>(crhs, nev, csink, t_enter, t_next, x0c, one, P.max_step, P.rtol, P.atol, r2, n_attempts);
                ++n_calls;
                tracked = r2.n_emitted > 0;
                S.n_samples[s] += r2.n_emitted;
            }
        }
        if (!tracked) S.closure[s] = CLOSURE_IMPOSSIBLE_OBS_TRACKING;
    }
    // Either the site reached index_t, or it is short of it and slides along the obstacle (:556-574)
    const int idx1 = site_index_t(S, s);
    if (idx1 < index_t) {
        const bool trigo = S.trigo[s] != 0;
        const int j0 = idx1 - idx0;
        const EvalGrid grid = make_eval_grid(t_start, t_end, n, j0 + 1);
        const Sample ys = S.xp[S.at(s, idx1)];
        const double x0c[2] = {ys.x0, ys.x1};
        ConstrRhs<FK> crhs{P, obs, trigo};
        QuitEvent<FK> qev{P, obs, false};
        ConstrSink<FK> csink{P, S, s, idx0, obs, trigo};
        IvpResult<2> r3;
        solve_ivp<2, METHOD>(crhs, qev, csink, grid.at(j0), grid.at(n - 1), x0c, grid, P.max_step, P.rtol, P.atol, r3,
                     n_attempts);
        ++n_calls;
        S.ivp_status[s] = r3.status;
        S.n_samples[s] += r3.n_emitted;
        if (qev.fired) S.closure[s] = CLOSURE_IMPOSSIBLE_OBS_TRACKING;
    }
}

// ---------------------------------------------------------------------------------------------------------
// NavigationProblem.auto_time_upper_bound (problem.py:193-239): two feedback flights from x_init, each integrated by
// solve_ivp (RK45, default tolerances, no step limit) until the target disc is entered -
//   mode 0: GSTargetFB (feedback.py:58-89)  - ground velocity aimed at the target, full airspeed;
//   mode 1: HTargetFB (feedback.py:92-130)  - heading aimed at the target, a UNIT control (not scaled by srf, as written).
// The result is the root of the terminal target event, +inf when the flight runs out of time (rel_timeout).
#define DABRY_EARTH_RADIUS 6378137.  // misc.py:183

template <int FK>
struct FeedbackRhs {
    const ProblemDev& P;
    int mode;
    DABRY_HD void operator()(double t, const double* x, double* dx) const {
        double e0, e1;
        if (P.coords == COORDS_GCS) {  // chord to the target in the local east / north frame (feedback.py:69-79, 114-123)
            double sl, cl, sp, cp, slt, clt, spt, cpt;
            sincos(x[0], &sl, &cl);
            sincos(x[1], &sp, &cp);
            sincos(P.x_target[0], &slt, &clt);
            sincos(P.x_target[1], &spt, &cpt);
            const double d0 = DABRY_EARTH_RADIUS * (clt * cpt) - DABRY_EARTH_RADIUS * (cl * cp);
            const double d1 = DABRY_EARTH_RADIUS * (slt * cpt) - DABRY_EARTH_RADIUS * (sl * cp);
            const double d2 = DABRY_EARTH_RADIUS * spt - DABRY_EARTH_RADIUS * sp;
            e0 = (d0 * -sl + d1 * cl) + d2 * 0.;
            e1 = (d0 * (-sp * cl) + d1 * (-sp * sl)) + d2 * cp;
        } else {
            e0 = P.x_target[0] - x[0];
            e1 = P.x_target[1] - x[1];
        }
        const double en = sqrt(e0 * e0 + e1 * e1);
        Flow f;
        field_eval<FK, false>(P.field, t, x[0], x[1], f);
        double u0 = 0., u1 = 0.;
        if (!(en <= 1e-8)) {  // np.isclose(norm, 0)
            if (mode == 0) {
                directional_control(f.w0, f.w1, e0, e1, P.srf, u0, u1);
            } else {
                const double ang = atan2(e1 / en, e0 / en);
                sincos(ang, &u1, &u0);
            }
        }
        if (P.coords == COORDS_GCS) {
            dx[0] = (1. / cos(x[1])) * (u0 + f.w0);
            dx[1] = u1 + f.w1;
        } else {
            dx[0] = u0 + f.w0;
            dx[1] = u1 + f.w1;
        }
    }
};

struct TargetReached {  // event_target of apply_feedback (problem.py:209-212): terminal, no direction attribute = both
    static constexpr bool kBothDirections = true;
    const ProblemDev& P;
    DABRY_HD int count() const { return 1; }
    DABRY_HD bool terminal(int) const { return true; }
    DABRY_HD double value(int, double, const double* x) const { return event_value(P, 0, x[0], x[1]); }
    DABRY_HD void record(int, double) {}
};

struct NoSink {
    DABRY_HD void emit(int, double, const double*) {}
};

template <int FK>
DABRY_HD double feedback_flight_time(const ProblemDev& P, int mode, double x0, double x1, double t_start, double t_end,
                                     unsigned& n_attempts) {
    FeedbackRhs<FK> rhs{P, mode};
    TargetReached ev{P};
    NoSink sink;
    const EvalGrid none = make_eval_grid(t_start, t_end, 0, 0);  // t_eval only samples: it does not steer the steps
    IvpResult<2> res;
    const double y0[2] = {x0, x1};
    solve_ivp<2, METHOD_DOPRI5>(rhs, ev, sink, t_start, t_end, y0, none, INFINITY, P.rtol, P.atol, res, n_attempts);
    return res.status == IVP_TERMINATED ? res.t_term : INFINITY;
}

// ---------------------------------------------------------------------------------------------------------
// connect_to_parents (solver_ef.py:690-699)
DABRY_HD void site_connect(const SiteStore& S, int s) {
    const int a = S.parent_prev[s], b = S.parent_next[s];
    if (a < 0 || b < 0 || (S.flags[s] & SITE_CONNECTED)) return;  // roots; `not site.has_neighbours` (:809)
    S.flags[s] |= SITE_CONNECTED;
    const int k = S.index_t_init[s];
    S.next_nb[S.at(a, k)] = s;
    S.next_nb[S.at(s, k)] = b;
    S.check_next[a] = k;
}

// compute_new_sites (solver_ef.py:630-665) for one site.  Returns the grid index at which a child must be
// created between the site and its neighbour (-1: none).  The child itself is built by site_spawn once a
// prefix sum has assigned its slot, which reproduces the insertion order of the reference's dict.
DABRY_HD int site_resample(const ProblemDev& P, const SolverParams& C, const SiteStore& S, int s,
                           int index_t_hi) {
    const int k0 = S.check_next[s];
    const int nb = S.next_nb[S.at(s, k0)];
    if (nb < 0) return -1;
    int hi = imin(imin(site_index_t(S, s) + 1, site_index_t(S, nb) + 1), index_t_hi);
    int new_check = k0;
    int request = -1;
    for (int k = k0; k < hi; ++k) {
        const bool a_in = site_in_obs_at(S, s, k), b_in = site_in_obs_at(S, nb, k);
        if (a_in && b_in) {
            S.neuter[s] = NEUTER_SELF_AND_NB_IN_OBS;
            S.index_neutered[s] = k;
            break;
        }
        const Sample a = S.xp[S.at(s, k)], b = S.xp[S.at(nb, k)];
        const double d0 = a.x0 - b.x0, d1 = a.x1 - b.x1;
        if (add_rn(mul_rn(d0, d0), mul_rn(d1, d1)) > C.max_dist_sq) {
            const int index = (C.mode_origin && !a_in && !b_in && S.index_t_init[s] == 0 &&
                               S.index_t_init[nb] == 0) ? 0 : k;
            if (S.depth[s] < C.max_depth - 1 && S.depth[nb] < C.max_depth - 1) {
                const Sample ca = S.xp[S.at(s, index)], cb = S.xp[S.at(nb, index)];
                const double c0 = 0.5 * (ca.x0 + cb.x0), c1 = 0.5 * (ca.x1 + cb.x1);
                if (in_any_obstacle(P, c0, c1)) S.closure[s] = CLOSURE_WITHIN_OBS;
                else request = index;
            }
            break;
        }
        new_check = k;
    }
    for (int k = k0; k <= new_check; ++k) S.next_nb[S.at(s, k)] = nb;
    S.check_next[s] = new_check;
    return request;
}

// SiteManager.site_from_parents (solver_ef.py:261-278) into slot `child`
DABRY_HD void site_spawn(const SolverParams& C, const SiteStore& S, int a, int child, int index, int subframe) {
    const int b = S.next_nb[S.at(a, S.check_next[a])];
    site_reset(S, child);
    const int64_t ra = S.at(a, index), rb = S.at(b, index);
    const Sample sa = S.xp[ra], sb = S.xp[rb];
    double pa0 = sa.p0, pa1 = sa.p1, pb0 = sb.p0, pb1 = sb.p1;
    if (site_in_obs_at(S, a, index)) {
        const double nrm = sqrt(pb0 * pb0 + pb1 * pb1);
        pa0 = -S.ctrl[ra].u0 * nrm;
        pa1 = -S.ctrl[ra].u1 * nrm;
    }
    if (site_in_obs_at(S, b, index)) {
        const double nrm = sqrt(pa0 * pa0 + pa1 * pa1);
        pb0 = -S.ctrl[rb].u0 * nrm;
        pb1 = -S.ctrl[rb].u1 * nrm;
    }
    const int64_t rc = S.at(child, index);
    S.xp[rc] = Sample{0.5 * (sa.x0 + sb.x0), 0.5 * (sa.x1 + sb.x1), 0.5 * (pa0 + pb0), 0.5 * (pa1 + pb1)};
    S.ctrl[rc] = Control{quiet_nan(), quiet_nan()};
    S.tt[rc] = S.tt[ra];
    S.cost0[child] = 0.5 * (site_cost_at(S, a, index) + site_cost_at(S, b, index));
    S.index_t_init[child] = index;
    arrival_note(problem(), S, child, index, 0.5 * (sa.x0 + sb.x0), 0.5 * (sa.x1 + sb.x1));  // the child's first sample
    S.check_next[child] = index;
    S.parent_prev[child] = a;
    S.parent_next[child] = b;
    S.created_subframe[child] = subframe;
    // dyadic index of the midpoint, with ring wrap (solver_ef.py:238-251)
    const int64_t n_total = (int64_t) C.n_roots << C.max_depth;
    const int64_t ia = S.dyadic[a];
    int64_t ib = S.dyadic[b];
    if (ib < ia) ib += n_total;
    const int64_t ic = (ia + ib) / 2;
    S.dyadic[child] = ic;
    const int64_t u = ic & (((int64_t) 1 << C.max_depth) - 1);
    int val = 0;
    if (u != 0) while (((u >> val) & 1) == 0) ++val;
    S.depth[child] = u == 0 ? 0 : C.max_depth - val;
}

// trim_distance (solver_ef.py:594-600)
DABRY_HD void site_trim_distance(const ProblemDev& P, const SolverParams& C, const SiteStore& S, int s) {
    const int k = site_index_t(S, s);
    const Sample a = S.xp[S.at(s, k)];
    const double d0 = a.x0 - P.x_target[0], d1 = a.x1 - P.x_target[1];
    const double sup_time = sqrt(d0 * d0 + d1 * d1) / C.ff_max_norm;
    if (S.times[S.n_time - 1] - S.times[k] > sup_time) S.closure[s] = CLOSURE_SUBOPTIMAL;
}

// check_solutions (solver_ef.py:667-685): smallest cost among the grid samples strictly inside the target disc.
// Returns the grid index of that sample (-1: the site never enters the disc) and its cost.
DABRY_HD int site_arrival(const ProblemDev& P, const SiteStore& S, int s, double& cost) {
    const int k0 = S.index_t_init[s], n = S.n_samples[s];
    int best = -1;
    double best_cost = 0.;
    for (int k = k0; k < k0 + n; ++k) {
        const Sample a = S.xp[S.at(s, k)];
        const double d0 = a.x0 - P.x_target[0], d1 = a.x1 - P.x_target[1];
        if (add_rn(mul_rn(d0, d0), mul_rn(d1, d1)) < P.target_radius_sq) {
            const double c = site_cost_at(S, s, k);
            if (best < 0 || c < best_cost) {
                best = k;
                best_cost = c;
            }
        }
    }
    cost = best_cost;
    return best;
}

// ---------------------------------------------------------------------------------------------------------
// Trimming: cost map by triangle rasterisation (solver_ef.py:860-889, misc.py:115-139)

DABRY_HD void atomic_min_double(double* addr, double v) {
#if defined(__CUDA_ARCH__)
    // costs are >= 0 and the map starts at +inf: the IEEE bit patterns order like unsigned integers
    atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long) __double_as_longlong(v));
#else
    if (v < *addr) *addr = v;
#endif
}

DABRY_HD double cross2(double a0, double a1, double b0, double b1) {  // np.cross on 2-vectors
    return sub_rn(mul_rn(a0, b1), mul_rn(a1, b0));
}

// linspace(bl, tr, n)[i]
DABRY_HD double map_coord(double lo, double hi, int n, int i) {
    if (i == n - 1) return hi;
    return add_rn(mul_rn((double) i, (hi - lo) / (double) (n - 1)), lo);
}

DABRY_HD void raster_triangle(const SolverParams& C, double* map, double p10, double p11, double p20, double p21,
                              double p30, double p31, double c1, double c2, double c3) {
    const double v10 = sub_rn(p20, p10), v11 = sub_rn(p21, p11);
    const double v20 = sub_rn(p30, p10), v21 = sub_rn(p31, p11);
    const double den = cross2(v10, v11, v20, v21);
    if (!(fabs(den) > 1e-8)) return;  // np.isclose(cross, 0.): |c| <= atol = 1e-8  (NaN vertices: nothing inside)
    if (den != den) return;
    // bounding box in map indices, one cell of margin on each side
    const double xmin = dmin(p10, dmin(p20, p30)), xmax = dmax(p10, dmax(p20, p30));
    const double ymin = dmin(p11, dmin(p21, p31)), ymax = dmax(p11, dmax(p21, p31));
    if (!(xmin == xmin) || !(xmax == xmax) || !(ymin == ymin) || !(ymax == ymax)) return;
    // grid nodes that can lie inside: linspace index range of the bounding box, widened by 1e-6 of a cell against
    // the rounding of the index mapping (the inside test below is the reference's, evaluated per node; the reciprocal
    // cell size is one rounding, 1e-16, inside that margin).  On a dense front almost every triangle is smaller than a
    // cell and the range is empty.
    const double ax = (xmin - C.bl[0]) * C.cm_inv[0], bx = (xmax - C.bl[0]) * C.cm_inv[0];
    const double ay = (ymin - C.bl[1]) * C.cm_inv[1], by = (ymax - C.bl[1]) * C.cm_inv[1];
    if (!(bx >= -1. && ax <= (double) C.cm_nx && by >= -1. && ay <= (double) C.cm_ny)) return;
    const int i_lo = imax((int) ceil(ax - 1e-6), 0), i_hi = imin((int) floor(bx + 1e-6), C.cm_nx - 1);
    const int j_lo = imax((int) ceil(ay - 1e-6), 0), j_hi = imin((int) floor(by + 1e-6), C.cm_ny - 1);
    if (i_lo > i_hi || j_lo > j_hi) return;
    const double c02 = cross2(p10, p11, v20, v21);
    const double c01 = cross2(p10, p11, v10, v11);
    for (int i = i_lo; i <= i_hi; ++i) {
        const double gx = map_coord(C.bl[0], C.tr[0], C.cm_nx, i);
        for (int j = j_lo; j <= j_hi; ++j) {
            const double gy = map_coord(C.bl[1], C.tr[1], C.cm_ny, j);
            const double a = sub_rn(cross2(gx, gy, v20, v21), c02) / den;
            const double b = -(sub_rn(cross2(gx, gy, v10, v11), c01) / den);
            if (a > 0. && b > 0. && add_rn(a, b) < 1.) {
                const double ab = add_rn(a, b);
                const double cost = add_rn(add_rn(mul_rn(sub_rn(1., ab), c1), mul_rn(a, c2)), mul_rn(b, c3));
                atomic_min_double(map + (int64_t) i * C.cm_ny + j, cost);
            }
        }
    }
}

// one (site, grid index) cell of the loop at solver_ef.py:866-888
DABRY_HD void site_rasterize(const SolverParams& C, const SiteStore& S, double* map, int s, int k) {
    const int nb = S.next_nb[S.at(s, k)];
    if (nb < 0) return;
    const Sample a0 = S.xp[S.at(s, k)], a1 = S.xp[S.at(s, k + 1)];
    const Sample b0 = S.xp[S.at(nb, k)], b1 = S.xp[S.at(nb, k + 1)];
    const double c1 = site_cost_at(S, s, k), c3 = site_cost_at(S, s, k + 1);
    const double c2 = site_cost_at(S, nb, k), c4 = site_cost_at(S, nb, k + 1);
    raster_triangle(C, map, a0.x0, a0.x1, a1.x0, a1.x1, b1.x0, b1.x1, c1, c3, c4);
    if (k > 0) raster_triangle(C, map, a0.x0, a0.x1, b1.x0, b1.x1, b0.x0, b0.x1, c1, c4, c2);
}

// Utils.interpolate on the 2-D cost map (misc.py:570-589 as called at solver_ef.py:832)
DABRY_HD double cost_map_lookup(const SolverParams& C, const double* map, double x0, double x1) {
    const double sx = (C.tr[0] - C.bl[0]) / (double) (C.cm_nx - 1), sy = (C.tr[1] - C.bl[1]) / (double) (C.cm_ny - 1);
    const double px = (x0 - C.bl[0]) / sx, py = (x1 - C.bl[1]) / sy;
    const double fx = floor(px), fy = floor(py);
    const double wxh = px - fx, wyh = py - fy, wxl = 1. - wxh, wyl = 1. - wyh;
    const int i0 = iclamp((int) fx, 0, C.cm_nx - 1), i1 = iclamp((int) fx + 1, 0, C.cm_nx - 1);
    const int j0 = iclamp((int) fy, 0, C.cm_ny - 1), j1 = iclamp((int) fy + 1, 0, C.cm_ny - 1);
    const double v00 = map[(int64_t) i0 * C.cm_ny + j0], v01 = map[(int64_t) i0 * C.cm_ny + j1];
    const double v10 = map[(int64_t) i1 * C.cm_ny + j0], v11 = map[(int64_t) i1 * C.cm_ny + j1];
    return add_rn(add_rn(add_rn(mul_rn(mul_rn(wxl, wyl), v00), mul_rn(mul_rn(wxl, wyh), v01)),
                         mul_rn(mul_rn(wxh, wyl), v10)), mul_rn(mul_rn(wxh, wyh), v11));
}

// SolverEFTrimming.trim for one open site of sites_considered (solver_ef.py:831-839); true = still valid
DABRY_HD bool site_trim(const SolverParams& C, const SiteStore& S, const double* map, int s, int index) {
    const Sample a = S.xp[S.at(s, index)];
    const double v = cost_map_lookup(C, map, a.x0, a.x1);
    if (v != v) return true;
    if (site_cost_at(S, s, index) > 1.02 * v) {
        S.closure[s] = CLOSURE_SUBOPTIMAL;
        return false;
    }
    return true;
}

}  // namespace dabry
