// scipy.integrate.RK45 restated for one extremal per thread (scipy 1.18.1, scipy/integrate/_ivp):
//   rk.py:14-71 rk_step, :111-176 _step_impl (controller), :538-565 tableau, :715-737 dense output,
//   common.py:68-134 select_initial_step, :63-65 RMS norm;  scipy/optimize/Zeros/brentq.c for event roots.
// The reference calls solve_ivp with the default method RK45, rtol 1e-3, atol 1e-6 and max_step
// (solver_ef.py:427-429, 494-496, 528-532, 557-561).
#pragma once
#include "common.cuh"

namespace dabry {

#define DABRY_RK_SAFETY 0.9
#define DABRY_RK_MIN_FACTOR 0.2
#define DABRY_RK_MAX_FACTOR 10.0

template <int N>
struct RkState {
    double t, t_old, h_abs;
    double y[N], y_old[N], f[N];
    double K[7][N];  // K[0] = f(t_old, y_old) ... K[6] = f(t, y)   (rk.py:61-69)
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

// RungeKutta.__init__ (rk.py:86-103): f0 and the empirical first step.  Costs 2 RHS evaluations.
template <int N, class Rhs>
DABRY_HD void rk_init(RkState<N>& S, const Rhs& rhs, double t0, const double* y0, double t_bound, double max_step,
                      double rtol, double atol) {
    S.t = t0;
    S.t_old = t0;
#pragma unroll
    for (int i = 0; i < N; ++i) S.y[i] = S.y_old[i] = y0[i];
    rhs(t0, S.y, S.f);
    // select_initial_step (common.py:68-134), direction = +1, error_estimator_order = 4
    const double interval = fabs(t_bound - t0);
    if (interval == 0.) {
        S.h_abs = 0.;
        return;
    }
    double scale[N];
#pragma unroll
    for (int i = 0; i < N; ++i) scale[i] = atol + fabs(S.y[i]) * rtol;
    const double d0 = rms_norm_scaled<N>(S.y, scale);
    const double d1 = rms_norm_scaled<N>(S.f, scale);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    h0 = dmin(h0, interval);
    double y1[N], f1[N];
#pragma unroll
    for (int i = 0; i < N; ++i) y1[i] = S.y[i] + h0 * S.f[i];
    rhs(t0 + h0, y1, f1);
#pragma unroll
    for (int i = 0; i < N; ++i) f1[i] -= S.f[i];
    const double d2 = rms_norm_scaled<N>(f1, scale) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = dmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / dmax(d1, d2), 1. / 5.);
    S.h_abs = dmin(dmin(100 * h0, h1), dmin(interval, max_step));
}

// One accepted step of RungeKutta._step_impl (rk.py:111-176).  Returns false on TOO_SMALL_STEP.
// n_attempts is incremented once per attempted step (accepted or rejected): the metric's unit.
template <int N, class Rhs>
DABRY_HD bool rk_step(RkState<N>& S, const Rhs& rhs, double t_bound, double max_step, double rtol, double atol,
                      unsigned& n_attempts) {
    const double t = S.t;
    const double min_step = 10. * ulp_above(t);
    double h_abs = S.h_abs;
    if (h_abs > max_step) h_abs = max_step;
    else if (h_abs < min_step) h_abs = min_step;
    bool rejected = false;
    double y_new[N];
#pragma unroll
    for (int i = 0; i < N; ++i) S.K[0][i] = S.f[i];
    for (;;) {
        if (h_abs < min_step) return false;
        double t_new = t + h_abs;
        if (t_new - t_bound > 0.) t_new = t_bound;
        const double h = t_new - t;
        h_abs = fabs(h);
        ++n_attempts;
        // Dormand-Prince stages (rk.py:538-552)
        double ys[N];
#pragma unroll
        for (int i = 0; i < N; ++i) ys[i] = S.y[i] + (1. / 5. * S.K[0][i]) * h;
        rhs(t + 1. / 5. * h, ys, S.K[1]);
#pragma unroll
        for (int i = 0; i < N; ++i) ys[i] = S.y[i] + (3. / 40. * S.K[0][i] + 9. / 40. * S.K[1][i]) * h;
        rhs(t + 3. / 10. * h, ys, S.K[2]);
#pragma unroll
        for (int i = 0; i < N; ++i)
            ys[i] = S.y[i] + (44. / 45. * S.K[0][i] - 56. / 15. * S.K[1][i] + 32. / 9. * S.K[2][i]) * h;
        rhs(t + 4. / 5. * h, ys, S.K[3]);
#pragma unroll
        for (int i = 0; i < N; ++i)
            ys[i] = S.y[i] + (19372. / 6561. * S.K[0][i] - 25360. / 2187. * S.K[1][i] + 64448. / 6561. * S.K[2][i]
                              - 212. / 729. * S.K[3][i]) * h;
        rhs(t + 8. / 9. * h, ys, S.K[4]);
#pragma unroll
        for (int i = 0; i < N; ++i)
            ys[i] = S.y[i] + (9017. / 3168. * S.K[0][i] - 355. / 33. * S.K[1][i] + 46732. / 5247. * S.K[2][i]
                              + 49. / 176. * S.K[3][i] - 5103. / 18656. * S.K[4][i]) * h;
        rhs(t + h, ys, S.K[5]);
#pragma unroll
        for (int i = 0; i < N; ++i)
            y_new[i] = S.y[i] + h * (35. / 384. * S.K[0][i] + 500. / 1113. * S.K[2][i] + 125. / 192. * S.K[3][i]
                                     - 2187. / 6784. * S.K[4][i] + 11. / 84. * S.K[5][i]);
        rhs(t + h, y_new, S.K[6]);
        // error estimate (rk.py:105-109, :146-147)
        double err[N], scale[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            err[i] = (-71. / 57600. * S.K[0][i] + 71. / 16695. * S.K[2][i] - 71. / 1920. * S.K[3][i]
                      + 17253. / 339200. * S.K[4][i] - 22. / 525. * S.K[5][i] + 1. / 40. * S.K[6][i]) * h;
            scale[i] = atol + dmax(fabs(S.y[i]), fabs(y_new[i])) * rtol;
        }
        const double error_norm = rms_norm_scaled<N>(err, scale);
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
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        S.y_old[i] = S.y[i];
        S.y[i] = y_new[i];
        S.f[i] = S.K[6][i];
    }
    S.h_abs = h_abs;
    return true;
}

// Quartic dense output of the last accepted step (rk.py:178-180, 715-737): Q = K^T P, P from rk.py:553-565.
template <int N>
struct Dense {
    double t_old, h;
    double y_old[N];
    double Q[N][4];
};

template <int N>
DABRY_HD void rk_dense(const RkState<N>& S, Dense<N>& D) {
    D.t_old = S.t_old;
    D.h = S.t - S.t_old;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        D.y_old[i] = S.y_old[i];
        const double k0 = S.K[0][i], k2 = S.K[2][i], k3 = S.K[3][i], k4 = S.K[4][i], k5 = S.K[5][i], k6 = S.K[6][i];
        D.Q[i][0] = k0;
        D.Q[i][1] = -8048581381. / 2820520608. * k0 + 131558114200. / 32700410799. * k2
                    - 1754552775. / 470086768. * k3 + 127303824393. / 49829197408. * k4
                    - 282668133. / 205662961. * k5 + 40617522. / 29380423. * k6;
        D.Q[i][2] = 8663915743. / 2820520608. * k0 - 68118460800. / 10900136933. * k2
                    + 14199869525. / 1410260304. * k3 - 318862633887. / 49829197408. * k4
                    + 2019193451. / 616988883. * k5 - 110615467. / 29380423. * k6;
        D.Q[i][3] = -12715105075. / 11282082432. * k0 + 87487479700. / 32700410799. * k2
                    - 10690763975. / 1880347072. * k3 + 701980252875. / 199316789632. * k4
                    - 1453857185. / 822651844. * k5 + 69997945. / 29380423. * k6;
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
