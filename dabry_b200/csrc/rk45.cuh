// scipy.integrate.RK45 restated for one extremal per thread (scipy 1.18.1, scipy/integrate/_ivp):
//   rk.py:14-71 rk_step, :111-176 _step_impl (controller), :538-565 tableau, :715-737 dense output,
//   common.py:68-134 select_initial_step, :63-65 RMS norm;  scipy/optimize/Zeros/brentq.c for event roots.
// The reference calls solve_ivp with the default method RK45, rtol 1e-3, atol 1e-6 and max_step
// (solver_ef.py:427-429, 494-496, 528-532, 557-561).
#pragma once
#include "common.cuh"

namespace dabry {

#ifndef DABRY_TRACE_RK_ATTEMPT  // analysis hook of tests/hostsim (one line per step attempt); nothing in the product
#define DABRY_TRACE_RK_ATTEMPT(n, t, h, error_norm)
#endif

#define DABRY_RK_SAFETY 0.9
#define DABRY_RK_MIN_FACTOR 0.2
#define DABRY_RK_MAX_FACTOR 10.0

// Dormand-Prince 5(4) coefficients (rk.py:538-565).  On the device they live in constant memory so that an FMA reads
// its coefficient straight from the constant bank instead of materialising a 64-bit immediate in two moves.
struct DormandPrince {
    double a21, a31, a32, a41, a42, a43, a51, a52, a53, a54, a61, a62, a63, a64, a65, b1, b3, b4, b5, b6, c2, c3, c4, c5, e1, e3, e4, e5, e6, e7, p11, p13, p14, p15, p16, p17, p21, p23, p24, p25, p26, p27, p31, p33, p34, p35, p36, p37;
};
#define DABRY_DP_INIT {1. / 5., 3. / 40., 9. / 40., 44. / 45., -56. / 15., 32. / 9., 19372. / 6561., -25360. / 2187., 64448. / 6561., -212. / 729., 9017. / 3168., -355. / 33., 46732. / 5247., 49. / 176., -5103. / 18656., 35. / 384., 500. / 1113., 125. / 192., -2187. / 6784., 11. / 84., 1. / 5., 3. / 10., 4. / 5., 8. / 9., -71. / 57600., 71. / 16695., -71. / 1920., 17253. / 339200., -22. / 525., 1. / 40., -8048581381. / 2820520608., 131558114200. / 32700410799., -1754552775. / 470086768., 127303824393. / 49829197408., -282668133. / 205662961., 40617522. / 29380423., 8663915743. / 2820520608., -68118460800. / 10900136933., 14199869525. / 1410260304., -318862633887. / 49829197408., 2019193451. / 616988883., -110615467. / 29380423., -12715105075. / 11282082432., 87487479700. / 32700410799., -10690763975. / 1880347072., 701980252875. / 199316789632., -1453857185. / 822651844., 69997945. / 29380423.}
#if defined(__CUDACC__)
__constant__ DormandPrince c_dp = DABRY_DP_INIT;
#endif
static const DormandPrince h_dp = DABRY_DP_INIT;
#if defined(__CUDA_ARCH__)
#define DP(x) (c_dp.x)
#else
#define DP(x) (h_dp.x)
#endif

// Where the stage values of a step live.  Rows 0..6 = K[0] = f(t_old, y_old) ... K[6] = f(t, y) (rk.py:61-69, after an
// accepted step K[6] is the FSAL value the next step starts from), row 7 = y_old.
//  * KRegisters: a plain array (registers / local memory) - host build, N = 2 arcs, every kernel but the hot one;
//  * KShared (device): one 8-byte column per thread in shared memory, row stride = block size (conflict-free, constant
//    offsets after unrolling).  The 32 doubles are live across the six non-inlined right-hand sides of a step; in
//    registers they leave the right-hand side ~50 registers to work with (its 24 node loads then go out two or three at
//    a time), in shared memory they cost one LDS each when a stage combination needs them.
template <int N>
struct KRegisters {
    double v[8][N];
    DABRY_HD double* operator[](int a) { return v[a]; }
    DABRY_HD const double* operator[](int a) const { return v[a]; }
    DABRY_HD void load_row(int a, double* out) const {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = v[a][i];
    }
    DABRY_HD void store_row(int a, const double* in) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[a][i] = in[i];
    }
};

#if defined(__CUDACC__)
// Layout: tile[row][pair][thread] of 16-byte pairs (components 2 * pair, 2 * pair + 1), so that a row of a thread is
// N / 2 conflict-free 128-bit accesses (a quarter-warp covers 128 contiguous bytes).
template <int N, int THREADS>
struct KShared {
    double* base;  // (double*) &tile2[threadIdx.x], tile2 = double2[8 * (N / 2) * THREADS] in shared memory
    struct Row {
        double* p;
        __device__ double& operator[](int i) const { return p[(i >> 1) * (2 * THREADS) + (i & 1)]; }
    };
    __device__ Row operator[](int a) const { return Row{base + a * (N * THREADS)}; }
    __device__ void load_row(int a, double* out) const {
        const double2* q = reinterpret_cast<const double2*>(base + a * (N * THREADS));
#pragma unroll
        for (int k = 0; k < N / 2; ++k) {
            const double2 v = q[k * THREADS];
            out[2 * k] = v.x;
            out[2 * k + 1] = v.y;
        }
    }
    __device__ void store_row(int a, const double* in) const {
        double2* q = reinterpret_cast<double2*>(base + a * (N * THREADS));
#pragma unroll
        for (int k = 0; k < N / 2; ++k) q[k * THREADS] = make_double2(in[2 * k], in[2 * k + 1]);
    }
};
#endif

template <int N, class KS = KRegisters<N>>
struct RkState {
    double t, t_old, h_abs;
    double y[N];
    KS K;
    bool retry = false;  // DEFER builds: the last attempt was rejected, the next call of rk_step repeats the step
};

template <int N>
DABRY_HD double rms_norm_scaled(const double* v, const double* scale) {
    double s = 0.;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double q = v[i] / scale[i];
        s += q * q;
    }
    return sqrt(s) / sqrt((double) N);
}

template <int N, class KS>
DABRY_HD void put_stage(KS& K, int stage, const double* k) {
#pragma unroll
    for (int i = 0; i < N; ++i) K[stage][i] = k[i];
}

// RungeKutta.__init__ (rk.py:86-103): f0 and the empirical first step.  Costs 2 RHS evaluations.
template <int N, class KS, class Rhs>
DABRY_HD bool rk_init(RkState<N, KS>& S, const Rhs& rhs, double t0, const double* y0, double t_bound, double max_step,
                      double rtol, double atol, double t_data = INFINITY) {
    S.t = t0;
    S.t_old = t0;
#pragma unroll
    for (int i = 0; i < N; ++i) S.y[i] = S.K[7][i] = y0[i];
    double f0[N];
    rhs(t0, S.y, f0);
#pragma unroll
    for (int i = 0; i < N; ++i) S.K[6][i] = f0[i];
    // select_initial_step (common.py:68-134), direction = +1, error_estimator_order = 4
    const double interval = fabs(t_bound - t0);
    if (interval == 0.) {
        S.h_abs = 0.;
        return true;
    }
    double scale[N];
#pragma unroll
    for (int i = 0; i < N; ++i) scale[i] = atol + fabs(S.y[i]) * rtol;
    const double d0 = rms_norm_scaled<N>(S.y, scale);
    const double d1 = rms_norm_scaled<N>(f0, scale);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    h0 = dmin(h0, interval);
    // t_data: time up to which the caller can serve the right-hand side (streamed flow grid).  The probe below is the one
    // evaluation whose distance from t0 the step limit does not bound: if it falls beyond, the start is deferred.
    if (t0 + h0 > t_data) return false;
    double y1[N], f1[N];
#pragma unroll
    for (int i = 0; i < N; ++i) y1[i] = S.y[i] + h0 * f0[i];
    rhs(t0 + h0, y1, f1);
#pragma unroll
    for (int i = 0; i < N; ++i) f1[i] -= f0[i];
    const double d2 = rms_norm_scaled<N>(f1, scale) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = dmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / dmax(d1, d2), 1. / 5.);
    S.h_abs = dmin(dmin(100 * h0, h1), dmin(interval, max_step));
    return true;
}

// One accepted step of RungeKutta._step_impl (rk.py:111-176).  Returns 1, or 0 on TOO_SMALL_STEP.
// n_attempts is incremented once per attempted step (accepted or rejected): the metric's unit.
// DEFER: a rejected attempt returns 2 instead of looping (scipy's `while not step_accepted`), and the caller calls again;
// per-extremal arithmetic is identical, but the lanes of a warp whose attempt was accepted do not wait for the retry of
// the ones that rejected - those repeat their step in the warp's next round.
template <int N, bool DEFER = false, class KS, class Rhs>
DABRY_HD int rk_step(RkState<N, KS>& S, const Rhs& rhs, double t_bound, double max_step, double rtol, double atol,
                     unsigned& n_attempts) {
    const double t = S.t;
    // min_step = 10 ulp(t) (rk.py:119) only matters when the step is within a few ulps of t: 10 ulp(t) <= 4.5e-15 |t|,
    // so any h_abs above 1e-12 (|t| + 1) is provably larger and nextafter is skipped
    double h_abs = S.h_abs;
    const double safe = 1e-12 * (fabs(t) + 1.);
    bool rejected = DEFER && S.retry;
    if (!rejected) {
        if (h_abs > max_step) h_abs = max_step;
        else if (!(h_abs > safe) && h_abs < 10. * ulp_above(t)) h_abs = 10. * ulp_above(t);
    }
    double y_new[N];
    if (!rejected) {
#pragma unroll
        for (int i = 0; i < N; ++i) S.K[0][i] = S.K[6][i];  // FSAL: f(t, y) of the last accepted step
    }
    for (;;) {
        if (!(h_abs > safe) && h_abs < 10. * ulp_above(t)) return 0;
        double t_new = t + h_abs;
        if (t_new - t_bound > 0.) t_new = t_bound;
        const double h = t_new - t;
        h_abs = fabs(h);
        ++n_attempts;
        // Dormand-Prince stages (rk.py:538-552)
        double ys[N], ks[N];
#pragma unroll
        for (int i = 0; i < N; ++i) ys[i] = S.y[i] + (DP(a21) * S.K[0][i]) * h;
        rhs(t + DP(c2) * h, ys, ks);
        put_stage<N>(S.K, 1, ks);
#pragma unroll
        for (int i = 0; i < N; ++i) ys[i] = S.y[i] + (DP(a31) * S.K[0][i] + DP(a32) * S.K[1][i]) * h;
        rhs(t + DP(c3) * h, ys, ks);
        put_stage<N>(S.K, 2, ks);
#pragma unroll
        for (int i = 0; i < N; ++i)
            ys[i] = S.y[i] + (DP(a41) * S.K[0][i] + DP(a42) * S.K[1][i] + DP(a43) * S.K[2][i]) * h;
        rhs(t + DP(c4) * h, ys, ks);
        put_stage<N>(S.K, 3, ks);
#pragma unroll
        for (int i = 0; i < N; ++i)
            ys[i] = S.y[i] + (DP(a51) * S.K[0][i] + DP(a52) * S.K[1][i] + DP(a53) * S.K[2][i]
                              + DP(a54) * S.K[3][i]) * h;
        rhs(t + DP(c5) * h, ys, ks);
        put_stage<N>(S.K, 4, ks);
#pragma unroll
        for (int i = 0; i < N; ++i)
            ys[i] = S.y[i] + (DP(a61) * S.K[0][i] + DP(a62) * S.K[1][i] + DP(a63) * S.K[2][i]
                              + DP(a64) * S.K[3][i] + DP(a65) * S.K[4][i]) * h;
        rhs(t + h, ys, ks);
        put_stage<N>(S.K, 5, ks);
#pragma unroll
        for (int i = 0; i < N; ++i)
            y_new[i] = S.y[i] + h * (DP(b1) * S.K[0][i] + DP(b3) * S.K[2][i] + DP(b4) * S.K[3][i]
                                     + DP(b5) * S.K[4][i] + DP(b6) * S.K[5][i]);
        rhs(t + h, y_new, ks);
        put_stage<N>(S.K, 6, ks);
        // error estimate (rk.py:105-109, :146-147)
        double err[N], scale[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            err[i] = (DP(e1) * S.K[0][i] + DP(e3) * S.K[2][i] + DP(e4) * S.K[3][i]
                      + DP(e5) * S.K[4][i] + DP(e6) * S.K[5][i] + DP(e7) * S.K[6][i]) * h;
            scale[i] = atol + dmax(fabs(S.y[i]), fabs(y_new[i])) * rtol;
        }
        const double error_norm = rms_norm_scaled<N>(err, scale);
        DABRY_TRACE_RK_ATTEMPT(N, t, h, error_norm);
        if (error_norm < 1.) {
            // The next step starts from min(h_abs * factor, max_step) (rk.py:121-126), so the power is only needed
            // when h_abs * factor can land below max_step.  factor >= r  <=>  error_norm <= (0.9 / r)^5; the test
            // keeps a 1 % margin so that a borderline case still takes the exact path.  With h_abs == max_step
            // (~99 % of the steps of the reference's runs) this is error_norm < 0.58.
            const double r = max_step / h_abs;
            const double q = DABRY_RK_SAFETY / r;
            const double q2 = q * q;
            if (!rejected && r <= DABRY_RK_MAX_FACTOR && error_norm < 0.99 * (q2 * q2 * q)) {
                h_abs = max_step;
            } else {
                double factor;
                if (error_norm == 0.) factor = DABRY_RK_MAX_FACTOR;
                else factor = dmin(DABRY_RK_MAX_FACTOR, DABRY_RK_SAFETY * pow(error_norm, -0.2));
                if (rejected) factor = dmin(1., factor);
                h_abs *= factor;
            }
            S.t_old = t;
            S.t = t_new;
            break;
        }
        h_abs *= dmax(DABRY_RK_MIN_FACTOR, DABRY_RK_SAFETY * pow(error_norm, -0.2));
        rejected = true;
        if (DEFER) {  // K[0] = f(t, y) and y are untouched: the next call repeats the step with the smaller h_abs
            S.h_abs = h_abs;
            S.retry = true;
            return 2;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        S.K[7][i] = S.y[i];
        S.y[i] = y_new[i];
    }
    S.h_abs = h_abs;
    if (DEFER) S.retry = false;
    return 1;
}

// Coefficients of the local interpolant of a step (filled by rk_dense / rk4_dense / rk_step_rolled).
template <int N>
struct Dense {
    double t_old, h;
    double y_old[N];
    double Q[N][4];
};

// ---- rolled variant of the step (the throughput kernel of gridded fields, site.cuh DABRY_FAST_STEP) -------------------
// The six stage evaluations as ONE loop body: the right-hand side appears once in the step, so it can be inlined without
// multiplying the code by six (no call, no argument marshalling, its loads scheduled with the stage sums); the tableau
// row is read from constant memory by stage index, and the stage values are addressed by stage index - they live in
// shared memory (KShared: two 128-bit accesses per row) instead of registers, which the right-hand side then has to
// itself.  Row 6 of the table is b (FSAL: y_new is the sixth combination, f(t + h, y_new) the seventh stage).  Same sums
// in the same order as rk_step; zero coefficients stay in the sums, as numpy's dot keeps them.  Measured (B200, config 4):
// 24.7 vs 25.5 ms for the unrolled step with a non-inlined right-hand side, config 3 14.1 vs 16.2 ms
// (profiles/r02_ab_rolled.txt); the same loop with the stage values in registers behind a switch, or in local memory, or
// with a called right-hand side is slower than the unrolled step (32.2 / 27.9 / 27.1 ms).
struct DormandPrinceRows {
    double a[6][6];  // row s - 1: coefficients of K[0..s-1] for stage s = 1..5, row 5: b
    double c[6];
};
#define DABRY_DP_ROWS_INIT {{{1. / 5., 0., 0., 0., 0., 0.}, {3. / 40., 9. / 40., 0., 0., 0., 0.}, {44. / 45., -56. / 15., 32. / 9., 0., 0., 0.}, {19372. / 6561., -25360. / 2187., 64448. / 6561., -212. / 729., 0., 0.}, {9017. / 3168., -355. / 33., 46732. / 5247., 49. / 176., -5103. / 18656., 0.}, {35. / 384., 0., 500. / 1113., 125. / 192., -2187. / 6784., 11. / 84.}}, {1. / 5., 3. / 10., 4. / 5., 8. / 9., 1., 1.}}
#if defined(__CUDACC__)
__constant__ DormandPrinceRows c_dp_rows = DABRY_DP_ROWS_INIT;
#endif
static const DormandPrinceRows h_dp_rows = DABRY_DP_ROWS_INIT;
#if defined(__CUDA_ARCH__)
#define DPR (c_dp_rows)
#else
#define DPR (h_dp_rows)
#endif

// DABRY_FAST_KEEP: what the rolled step keeps in registers besides shared memory - 1: K[0] (read by every stage),
// 2: the row the last right-hand side produced (the last term of the next stage's sum), 4: the rows of the error estimate
// for the interpolant of the accepted step (built inside the step instead of by rk_dense).
// Measured (profiles/r02_ab_rolled.txt, free arcs of configs 4 / 4-fp32 / 3 / 5): 0: 24.70 / 21.29 / 14.06 / 17.45 ms,
// 1: 24.55 / 21.37 / 14.40 / 17.46, 2: 24.31 / 21.29 / 13.54 / 17.34, 3: 24.50 / 21.62 / 14.45 / 17.35, 4: 26.1 (config 4).
#ifndef DABRY_FAST_KEEP
#define DABRY_FAST_KEEP 2
#endif
#ifndef DABRY_FAST_UNROLL_J  // 1: the row loads of a stage unrolled and predicated instead of a loop over the rows
#define DABRY_FAST_UNROLL_J 0  // measured: configs 4 / 4-fp32 / 3 / 5 free arcs 24.55 / 21.05 / 13.43 / 17.17 ms vs 24.32 / 21.31 / 13.56 / 17.28
#endif
// dense: when not null, the interpolant of the accepted step is built here, from the rows the error estimate has just
// read (rk_dense would read the same six rows from shared memory again).
template <int N, class KS, class Rhs>
DABRY_HD int rk_step_rolled(RkState<N, KS>& S, const Rhs& rhs, double t_bound, double max_step, double rtol, double atol,
                            unsigned& n_attempts, Dense<N>* dense = nullptr) {
    const double t = S.t;
    double h_abs = S.h_abs;
    const double safe = 1e-12 * (fabs(t) + 1.);
    bool rejected = false;
    if (h_abs > max_step) h_abs = max_step;
    else if (!(h_abs > safe) && h_abs < 10. * ulp_above(t)) h_abs = 10. * ulp_above(t);
    double k0[N];  // FSAL: f(t, y) of the last accepted step
    S.K.load_row(6, k0);
    S.K.store_row(0, k0);
    double y_new[N];
    double q6[N], q2[N], q3[N], q4[N], q5[N];  // rows of the accepted attempt, for the interpolant
    for (;;) {
        if (!(h_abs > safe) && h_abs < 10. * ulp_above(t)) return 0;
        double t_new = t + h_abs;
        if (t_new - t_bound > 0.) t_new = t_bound;
        const double h = t_new - t;
        h_abs = fabs(h);
        ++n_attempts;
        double ks[N];  // the row the last right-hand side produced
#pragma unroll 1
        for (int s = 1; s <= 6; ++s) {
            double acc[N];
            {
                double r0[N];
                if (DABRY_FAST_KEEP & 1) {
#pragma unroll
                    for (int i = 0; i < N; ++i) r0[i] = k0[i];
                } else {
                    S.K.load_row(0, r0);
                }
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] = DPR.a[s - 1][0] * r0[i];
            }
            const int j_end = (DABRY_FAST_KEEP & 2) ? s - 1 : s;  // the last term comes from registers
#if DABRY_FAST_UNROLL_J
            // the row loads of a stage issued together, predicated on the stage index, instead of a loop over j
#pragma unroll
            for (int j = 1; j < 6; ++j) {
                if (j < j_end) {
                    const double a = DPR.a[s - 1][j];
                    double kj[N];
                    S.K.load_row(j, kj);
#pragma unroll
                    for (int i = 0; i < N; ++i) acc[i] = acc[i] + a * kj[i];
                }
            }
#else
#pragma unroll 1
            for (int j = 1; j < j_end; ++j) {
                const double a = DPR.a[s - 1][j];
                double kj[N];
                S.K.load_row(j, kj);
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] = acc[i] + a * kj[i];
            }
#endif
            if ((DABRY_FAST_KEEP & 2) && s >= 2) {
                const double a = DPR.a[s - 1][s - 1];
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] = acc[i] + a * ks[i];
            }
#pragma unroll
            for (int i = 0; i < N; ++i) y_new[i] = S.y[i] + acc[i] * h;
            rhs(s < 5 ? t + DPR.c[s - 1] * h : t + h, y_new, ks);
            S.K.store_row(s, ks);
        }
        double err[N], scale[N];
        {
            double r0[N];
            if (DABRY_FAST_KEEP & 1) {
#pragma unroll
                for (int i = 0; i < N; ++i) r0[i] = k0[i];
            } else {
                S.K.load_row(0, r0);
            }
            S.K.load_row(2, q2);
            S.K.load_row(3, q3);
            S.K.load_row(4, q4);
            S.K.load_row(5, q5);
            if (DABRY_FAST_KEEP & 2) {
#pragma unroll
                for (int i = 0; i < N; ++i) q6[i] = ks[i];
            } else {
                S.K.load_row(6, q6);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                err[i] = (DP(e1) * r0[i] + DP(e3) * q2[i] + DP(e4) * q3[i] + DP(e5) * q4[i] + DP(e6) * q5[i]
                          + DP(e7) * q6[i]) * h;
                scale[i] = atol + dmax(fabs(S.y[i]), fabs(y_new[i])) * rtol;
            }
        }
        const double error_norm = rms_norm_scaled<N>(err, scale);
        DABRY_TRACE_RK_ATTEMPT(N, t, h, error_norm);
        if (error_norm < 1.) {
            const double r = max_step / h_abs;
            const double q = DABRY_RK_SAFETY / r;
            const double q2 = q * q;
            if (!rejected && r <= DABRY_RK_MAX_FACTOR && error_norm < 0.99 * (q2 * q2 * q)) {
                h_abs = max_step;
            } else {
                double factor;
                if (error_norm == 0.) factor = DABRY_RK_MAX_FACTOR;
                else factor = dmin(DABRY_RK_MAX_FACTOR, DABRY_RK_SAFETY * pow(error_norm, -0.2));
                if (rejected) factor = dmin(1., factor);
                h_abs *= factor;
            }
            S.t_old = t;
            S.t = t_new;
            break;
        }
        h_abs *= dmax(DABRY_RK_MIN_FACTOR, DABRY_RK_SAFETY * pow(error_norm, -0.2));
        rejected = true;
    }
    if (dense) {  // rk_dense on the rows still in registers; y_old is the state the step started from
        dense->t_old = S.t_old;
        dense->h = S.t - S.t_old;
        double r0[N];
        if (DABRY_FAST_KEEP & 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) r0[i] = k0[i];
        } else {
            S.K.load_row(0, r0);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            dense->y_old[i] = S.y[i];
            const double c0 = r0[i], c2 = q2[i], c3 = q3[i], c4 = q4[i], c5 = q5[i], c6 = q6[i];
            dense->Q[i][0] = c0;
            dense->Q[i][1] = DP(p11) * c0 + DP(p13) * c2 + DP(p14) * c3 + DP(p15) * c4 + DP(p16) * c5 + DP(p17) * c6;
            dense->Q[i][2] = DP(p21) * c0 + DP(p23) * c2 + DP(p24) * c3 + DP(p25) * c4 + DP(p26) * c5 + DP(p27) * c6;
            dense->Q[i][3] = DP(p31) * c0 + DP(p33) * c2 + DP(p34) * c3 + DP(p35) * c4 + DP(p36) * c5 + DP(p37) * c6;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) S.y[i] = y_new[i];
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            S.K[7][i] = S.y[i];
            S.y[i] = y_new[i];
        }
    }
    S.h_abs = h_abs;
    return 1;
}

// Quartic dense output of the last accepted step (rk.py:178-180, 715-737): Q = K^T P, P from rk.py:553-565.

template <int N, class KS>
DABRY_HD void rk_dense(const RkState<N, KS>& S, Dense<N>& D) {
    D.t_old = S.t_old;
    D.h = S.t - S.t_old;
    double r0[N], r2[N], r3[N], r4[N], r5[N], r6[N], r7[N];
    S.K.load_row(0, r0);
    S.K.load_row(2, r2);
    S.K.load_row(3, r3);
    S.K.load_row(4, r4);
    S.K.load_row(5, r5);
    S.K.load_row(6, r6);
    S.K.load_row(7, r7);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        D.y_old[i] = r7[i];
        const double k0 = r0[i], k2 = r2[i], k3 = r3[i], k4 = r4[i], k5 = r5[i], k6 = r6[i];
        D.Q[i][0] = k0;
        D.Q[i][1] = DP(p11) * k0 + DP(p13) * k2 + DP(p14) * k3 + DP(p15) * k4 + DP(p16) * k5 + DP(p17) * k6;
        D.Q[i][2] = DP(p21) * k0 + DP(p23) * k2 + DP(p24) * k3 + DP(p25) * k4 + DP(p26) * k5 + DP(p27) * k6;
        D.Q[i][3] = DP(p31) * k0 + DP(p33) * k2 + DP(p34) * k3 + DP(p35) * k4 + DP(p36) * k5 + DP(p37) * k6;
    }
}

template <int N>
DABRY_HD void dense_eval(const Dense<N>& D, double t, double* y) {
    const double x = (t - D.t_old) / D.h;
    const double x2 = x * x, x3 = x2 * x, x4 = x3 * x;
#pragma unroll
    for (int i = 0; i < N; ++i)
        y[i] = D.h * (((D.Q[i][0] * x + D.Q[i][1] * x2) + D.Q[i][2] * x3) + D.Q[i][3] * x4) + D.y_old[i];
}

// ---- classic fixed-step RK4 (integrator RK4_FIXED) ---------------------------------------------------------------
// No reference counterpart (the reference only ever runs scipy's RK45): steps of exactly max_step (the last one clipped
// to t_bound), 4 right-hand sides per step (f(t_new, y_new) is kept as the next step's k1), cubic Hermite local
// interpolant for the time-grid samples and the event roots.  oracle/solver_port.py holds the same stepper as a
// scipy OdeSolver subclass, so the driver semantics (events, t_eval) are scipy's in both.
enum Integrator : int32_t { METHOD_DOPRI5 = 0, METHOD_RK4 = 1 };

template <int N, class KS, class Rhs>
DABRY_HD void rk4_init(RkState<N, KS>& S, const Rhs& rhs, double t0, const double* y0, double max_step) {
    S.t = t0;
    S.t_old = t0;
#pragma unroll
    for (int i = 0; i < N; ++i) S.y[i] = S.K[7][i] = y0[i];
    double f0[N];
    rhs(t0, S.y, f0);
    put_stage<N>(S.K, 6, f0);
    S.h_abs = max_step;
}

template <int N, class KS, class Rhs>
DABRY_HD bool rk4_step(RkState<N, KS>& S, const Rhs& rhs, double t_bound, double max_step, unsigned& n_attempts) {
    const double t = S.t;
    double t_new = t + max_step;
    if (t_new - t_bound > 0.) t_new = t_bound;
    const double h = t_new - t;
    ++n_attempts;
    double ys[N], k2[N], k3[N], k4[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        S.K[0][i] = S.K[6][i];
        ys[i] = S.y[i] + (0.5 * h) * S.K[0][i];
    }
    rhs(t + 0.5 * h, ys, k2);
#pragma unroll
    for (int i = 0; i < N; ++i) ys[i] = S.y[i] + (0.5 * h) * k2[i];
    rhs(t + 0.5 * h, ys, k3);
#pragma unroll
    for (int i = 0; i < N; ++i) ys[i] = S.y[i] + h * k3[i];
    rhs(t_new, ys, k4);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        S.K[7][i] = S.y[i];
        S.y[i] = S.y[i] + (h / 6.) * (((S.K[0][i] + 2. * k2[i]) + 2. * k3[i]) + k4[i]);
    }
    double f1[N];
    rhs(t_new, S.y, f1);
    put_stage<N>(S.K, 6, f1);
    S.t_old = t;
    S.t = t_new;
    return true;
}

// cubic Hermite interpolant in the polynomial form of Dense: y(t_old + x h) = y_old + h (Q0 x + Q1 x^2 + Q2 x^3)
template <int N, class KS>
DABRY_HD void rk4_dense(const RkState<N, KS>& S, Dense<N>& D) {
    D.t_old = S.t_old;
    D.h = S.t - S.t_old;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        D.y_old[i] = S.K[7][i];
        const double f0 = S.K[0][i], f1 = S.K[6][i];
        const double slope = (S.y[i] - D.y_old[i]) / D.h;
        D.Q[i][0] = f0;
        D.Q[i][1] = (3. * slope - 2. * f0) - f1;
        D.Q[i][2] = (f0 + f1) - 2. * slope;
        D.Q[i][3] = 0.;
    }
}

// scipy.optimize.brentq as solve_event_equation calls it (ivp.py:52-77): xtol = rtol = 4 eps, maxiter 100.
// Restated from the published algorithm (scipy/optimize/Zeros/brentq.c, Brent 1973 with inverse quadratic
// extrapolation); fa = f(xa), fb = f(xb) are evaluated by the routine itself like the C code does.
template <class F>
DABRY_HD_NOINLINE double brentq(const F& f, double xa, double xb) {
    const double xtol = 4. * 2.220446049250313e-16, rtol = 4. * 2.220446049250313e-16;
    double xpre = xa, xcur = xb, xblk = 0., fblk = 0., spre = 0., scur = 0.;
    double fpre = f(xpre), fcur = f(xcur);
    if (fpre == 0.) return xpre;
    if (fcur == 0.) return xcur;
    for (int it = 0; it < 100; ++it) {
        if (fpre != 0. && fcur != 0. && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur;
            xcur = xblk;
            xblk = xpre;
            fpre = fcur;
            fcur = fblk;
            fblk = fpre;
        }
        const double delta = (xtol + rtol * fabs(xcur)) / 2.;
        const double sbis = (xblk - xcur) / 2.;
        if (fcur == 0. || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);  // secant
            } else {
                const double dpre = (fpre - fcur) / (xpre - xcur);
                const double dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2. * fabs(stry) < dmin(fabs(spre), 3. * fabs(sbis) - delta)) {
                spre = scur;
                scur = stry;
            } else {
                spre = sbis;
                scur = sbis;
            }
        } else {
            spre = sbis;
            scur = sbis;
        }
        xpre = xcur;
        fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0. ? delta : -delta);
        fcur = f(xcur);
    }
    return xcur;
}

}  // namespace dabry
