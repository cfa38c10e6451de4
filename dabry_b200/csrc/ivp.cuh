// scipy.integrate.solve_ivp's driver loop restated for one extremal per thread (scipy 1.18.1,
// scipy/integrate/_ivp/ivp.py:659-732): step, event detection (find_active_events :134-158, all events of the
// reference have direction -1, misc.py:34-43), roots by brentq on the dense output (solve_event_equation
// :52-77), earliest terminal root truncates the step (handle_events :80-131), t_eval sampling through the
// dense output with searchsorted(side='right') (:711-728).
//
// The callers are the three solve_ivp call sites of the reference: solver_ef.py:427-429 / 494-496 (free arc,
// N = 4), :528-532 (one-sample constrained arc, N = 2, no event), :557-561 (constrained arc with quit event).
#pragma once
#include "rk45.cuh"

namespace dabry {

#define DABRY_MAX_EVENTS (1 + DABRY_MAX_OBSTACLES)

// DEFER builds (site.cuh DABRY_DEFER_RETRY): accepted steps a lane may run ahead of the slowest lane of its warp;
// negative = unbounded
#ifndef DABRY_DEFER_LEAD
#define DABRY_DEFER_LEAD -1
#endif

enum IvpStatus : int32_t { IVP_FINISHED = 0, IVP_TERMINATED = 1, IVP_FAILED = -1, IVP_RUNNING = 2, IVP_DEFERRED = 3 };

// Everything solve_ivp carries from one step to the next.  With a step budget the driver loop can be left after any
// step and re-entered later with bit-identical results, which lets the engine re-pack the still-running extremals into
// full warps between chunks of steps (extremals captured by an obstacle leave early).
template <int N>
struct IvpCarry {
    int32_t started;  // 0: fresh call
    int32_t eval_i;
    double t, h_abs;
    double y[N], f[N];
    double g[1 + DABRY_MAX_OBSTACLES];
};

// t_eval = numpy.linspace(t_start, t_end, n)[first:] (solver_ef.py:488, 495): j * step + t_start with the last
// element overwritten by t_end (numpy/_core/function_base.py linspace); products and sums individually rounded.
struct EvalGrid {
    double t_start, t_end, step;
    int n, first;
    DABRY_HD double at(int j) const {  // j indexes the un-sliced linspace
        return j == n - 1 ? t_end : add_rn(mul_rn((double) j, step), t_start);
    }
};

DABRY_HD EvalGrid make_eval_grid(double t_start, double t_end, int n, int first) {
    EvalGrid g;
    g.t_start = t_start;
    g.t_end = t_end;
    g.n = n;
    g.first = first;
    g.step = n > 1 ? (t_end - t_start) / (double) (n - 1) : 0.;
    return g;
}

template <int N>
struct IvpResult {
    int32_t status;
    int32_t n_emitted;       // number of t_eval samples produced
    int32_t term_event;      // index of the terminal event that stopped the integration, -1 if none
    double t_term;           // its root
    double y_term[N];        // sol(t_term)
};

// What the rare half of a step hands back: a step in which at least one event changed sign (handle_events + the t_eval
// samples up to the truncation point, ivp.py:80-131, 694-728).
template <int N>
struct EventStep {
    int32_t eval_i, n_emitted, term_event;  // term_event -1: no terminal event
    double t_term;
    double y_term[N];
};

template <int N, class Events, class Sink>
DABRY_HD_NOINLINE EventStep<N> event_step(const Dense<N> D, Events& events, Sink& sink, const EvalGrid& grid,
                                          int* act, int na, double t_old, double t, int eval_i) {
    EventStep<N> out;
    out.n_emitted = 0;
    out.term_event = -1;
    out.t_term = 0.;
    double roots[DABRY_MAX_EVENTS];
    bool any_terminal = false;
    for (int a = 0; a < na; ++a) {
        const int e = act[a];
        roots[a] = brentq([&](double tt) {
            double yy[N];
            dense_eval<N>(D, tt, yy);
            return events.value(e, tt, yy);
        }, t_old, t);
        any_terminal = any_terminal || events.terminal(e);
    }
    int keep = na;
    if (any_terminal) {
        // argsort(roots) on a tiny array (insertion sort, stable), cut after the first terminal
        for (int a = 1; a < na; ++a) {
            const double r = roots[a];
            const int e = act[a];
            int b = a - 1;
            while (b >= 0 && roots[b] > r) {
                roots[b + 1] = roots[b];
                act[b + 1] = act[b];
                --b;
            }
            roots[b + 1] = r;
            act[b + 1] = e;
        }
        for (int a = 0; a < na; ++a)
            if (events.terminal(act[a])) {
                keep = a + 1;
                break;
            }
    }
    for (int a = 0; a < keep; ++a) events.record(act[a], roots[a]);
    if (any_terminal) {
        out.term_event = act[keep - 1];
        t = roots[keep - 1];
        out.t_term = t;
        dense_eval<N>(D, t, out.y_term);
    }
    while (eval_i < grid.n) {  // t_eval values <= t (the truncated t) that were not emitted yet
        const double te = grid.at(eval_i);
        if (!(te <= t)) break;
        double yy[N];
        dense_eval<N>(D, te, yy);
        sink.emit(eval_i, te, yy);
        ++eval_i;
        ++out.n_emitted;
    }
    out.eval_i = eval_i;
    return out;
}

// Events concept:  int count() const;  bool terminal(int e) const;  double value(int e, double t, const double* y) const;
//                  void record(int e, double t_root);
// Sink concept:    void emit(int j, double t, const double* y);     j indexes the un-sliced linspace
template <int N, int METHOD, class KS = KRegisters<N>, bool DEFER = false, bool ROLLED = false, class Rhs, class Events,
          class Sink>
DABRY_HD void solve_ivp(const Rhs& rhs, Events& events, Sink& sink, double t0, double tf, const double* y0,
                        const EvalGrid& grid, double max_step, double rtol, double atol, IvpResult<N>& res,
                        unsigned& n_attempts, IvpCarry<N>* carry = nullptr, int step_budget = 0, KS kstore = KS(),
                        double t_park = INFINITY) {
    RkState<N, KS> S;
    S.K = kstore;
    const int ne = events.count();
    double g[DABRY_MAX_EVENTS];
    int eval_i;
    if (carry && carry->started) {
        S.t = S.t_old = carry->t;
        S.h_abs = carry->h_abs;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            S.y[i] = S.K[7][i] = carry->y[i];
            S.K[6][i] = carry->f[i];
        }
        for (int e = 0; e < ne; ++e) g[e] = carry->g[e];
        eval_i = carry->eval_i;
    } else {
        if (METHOD == METHOD_RK4) {
            rk4_init<N>(S, rhs, t0, y0, max_step);
        } else if (!rk_init<N>(S, rhs, t0, y0, tf, max_step, rtol, atol, carry ? t_park + max_step : INFINITY)) {
            res.status = IVP_DEFERRED;  // the initial-step probe needs flow data that has not arrived: start again later
            res.n_emitted = 0;
            return;
        }
        for (int e = 0; e < ne; ++e) g[e] = events.value(e, t0, y0);
        eval_i = grid.first;
    }
    int steps_left = step_budget;
    res.status = IVP_FINISHED;
    res.n_emitted = 0;
    int n_emitted = 0;  // written to res on the way out (res lives in the caller's frame)
    res.term_event = -1;
    res.t_term = 0.;
    bool finished = false;
#if defined(__CUDA_ARCH__) && DABRY_DEFER_LEAD >= 0
    // DEFER with a bounded lead: a lane may run at most DABRY_DEFER_LEAD accepted steps ahead of the slowest lane of its
    // warp.  Lead 0 is scipy's inner retry loop as the lockstep warp executes it (everybody waits for every retry),
    // an unbounded lead is plain DEFER (nobody waits, lanes drift apart by whole steps and the warp's gathers touch many
    // cells); in between, the lanes whose attempt was accepted use the retry rounds of the others for their own next
    // step.  The lanes that entered together vote at the top of every round; a lane that is done keeps voting until all are.
    const unsigned members = DEFER ? __activemask() : 0u;
    int steps_done = 0;
#endif
    for (;;) {
#if defined(__CUDA_ARCH__) && DABRY_DEFER_LEAD >= 0
        if (DEFER) {
            if (!__ballot_sync(members, !finished)) break;
#ifdef DABRY_DEFER_LEAD_T  // the lead measured in solver time (units of max_step) instead of accepted steps
            // warp minimum of a double over an arbitrary member mask: order-preserving map to an unsigned 64-bit key,
            // reduced in two 32-bit halves
            unsigned long long key = (unsigned long long) __double_as_longlong(finished ? INFINITY : S.t);
            key ^= (key >> 63) ? ~0ull : 0x8000000000000000ull;
            const unsigned hi = (unsigned) (key >> 32), lo = (unsigned) key;
            const unsigned hi_min = __reduce_min_sync(members, hi);
            const unsigned lo_min = __reduce_min_sync(members, hi == hi_min ? lo : 0xffffffffu);
            unsigned long long key_min = ((unsigned long long) hi_min << 32) | lo_min;
            key_min ^= (key_min >> 63) ? 0x8000000000000000ull : ~0ull;
            const double t_slowest = __longlong_as_double((long long) key_min);
            if (finished || S.t > t_slowest + (DABRY_DEFER_LEAD_T) * max_step) continue;
#else
            const int slowest = __reduce_min_sync(members, finished ? 0x7fffffff : steps_done);
            if (finished || steps_done > slowest + DABRY_DEFER_LEAD) continue;
#endif
        } else
#endif
        if (finished) break;
        // budget spent, or the time reached up to which the caller can serve the right-hand side: park between two steps
        if (carry && ((step_budget > 0 && steps_left-- == 0) || S.t >= t_park)) {
            carry->started = 1;
            carry->eval_i = eval_i;
            carry->t = S.t;
            carry->h_abs = S.h_abs;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                carry->y[i] = S.y[i];
                carry->f[i] = S.K[6][i];
            }
            for (int e = 0; e < ne; ++e) carry->g[e] = g[e];
            res.status = IVP_RUNNING;
            res.n_emitted = n_emitted;
            return;
        }
        Dense<N> D;  // the rolled step builds the interpolant of the step it accepts itself
        bool have_dense = false;
        // OdeSolver.step (base.py): a solver already at t_bound finishes without stepping
        if (S.t == tf) {
            S.t_old = S.t;
            finished = true;
        } else {
            const int ok = METHOD == METHOD_RK4 ? (int) rk4_step<N>(S, rhs, tf, max_step, n_attempts)
                           : (ROLLED && !DEFER)
                               ? rk_step_rolled<N>(S, rhs, tf, max_step, rtol, atol, n_attempts,
                                                   (DABRY_FAST_KEEP & 4) ? &D : nullptr)
                               : rk_step<N, DEFER>(S, rhs, tf, max_step, rtol, atol, n_attempts);
            have_dense = (DABRY_FAST_KEEP & 4) && ROLLED && !DEFER && METHOD != METHOD_RK4 && ok == 1;
            if (ok == 0) {
                res.status = IVP_FAILED;
                res.n_emitted = n_emitted;
#if defined(__CUDA_ARCH__) && DABRY_DEFER_LEAD >= 0
                if (DEFER) {  // keeps voting until the warp is done
                    finished = true;
                    continue;
                }
#endif
                return;
            }
            if (DEFER && ok == 2) continue;  // rejected: the same step again in the next round, nothing else happened
#if defined(__CUDA_ARCH__) && DABRY_DEFER_LEAD >= 0
            ++steps_done;
#endif
            if (S.t - tf >= 0.) finished = true;
        }
        const double t_old = S.t_old;
        const double t = S.t;
        int act[DABRY_MAX_EVENTS];
        int na = 0;
        for (int e = 0; e < ne; ++e) {
            const double g_new = events.value(e, t, S.y);
            // find_active_events (ivp.py:134-158): direction -1, or 0 (either way) for the feedback flights
            if ((g[e] >= 0. && g_new <= 0.) || (Events::kBothDirections && g[e] <= 0. && g_new >= 0.)) act[na++] = e;
            g[e] = g_new;
        }
        // the local interpolant is needed for event roots and for the t_eval samples inside this step
        const bool sample_due = eval_i < grid.n && grid.at(eval_i) <= t;
        if (na == 0 && !sample_due) continue;
        if (!have_dense) {
            if (METHOD == METHOD_RK4) rk4_dense<N>(S, D);
            else rk_dense<N>(S, D);
        }
        if (na == 0) {
            // common path, no call in it: the interpolant stays in registers (with the root finder in the same scope it
            // was written to local memory and read back on every step: 34 LDL/STL per accepted step)
            while (eval_i < grid.n) {  // t_eval values <= t that were not emitted yet
                const double te = grid.at(eval_i);
                if (!(te <= t)) break;
                double yy[N];
                dense_eval<N>(D, te, yy);
                sink.emit(eval_i, te, yy);
                ++eval_i;
                ++n_emitted;
            }
        } else {
            // rare path, out of line: roots of the active events, truncation by the first terminal one, samples up to it
            const EventStep<N> r = event_step<N>(D, events, sink, grid, act, na, t_old, t, eval_i);
            eval_i = r.eval_i;
            n_emitted += r.n_emitted;
            if (r.term_event >= 0) {
                res.status = IVP_TERMINATED;
                res.term_event = r.term_event;
                res.t_term = r.t_term;
#pragma unroll
                for (int i = 0; i < N; ++i) res.y_term[i] = r.y_term[i];
                finished = true;
            }
        }
    }
    res.n_emitted = n_emitted;
}

}  // namespace dabry
