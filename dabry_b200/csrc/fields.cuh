// Flow field evaluation: value w(t, x) and Jacobian J[a][b] = d w_a / d x_b in ONE pass.
//
// Two families (reference: dabry/flowfield.py):
//  * FIELD_GRID    - DiscreteFF (flowfield.py:178-515): N-linear clipped interpolation of the node values AND of
//                    the node gradients on the same 8 (4 if steady) nodes (misc.py:570-589, flowfield.py:451-472);
//  * FIELD_PROGRAM - a flattened expression tree of analytic leaves (flowfield.py:66-92 composition,
//                    565-1134 leaves): value = sum_i coeff_i * leaf_i.
// Both are seen through the non-dimensionalising WrapperFF (flowfield.py:518-544); identity parameters
// (time_origin 0, scales 1, bl 0) reproduce an unwrapped field bit for bit.
//
// HBM layout of a grid ("pair records"): for every (k, i, j), k = -1 .. nt - 1, one 96-byte record
//     { (c_max(k,0), c_min(k+1,nt-1)) : c in u, v, du/dx, du/dy, dv/dx, dv/dy }
// i.e. the two frames a time lerp needs are interleaved, a node is six aligned 16-byte loads, and the two
// nodes (j, j+1) of a cell edge are one contiguous 192-byte run (y is the fastest grid axis, as in the
// reference's values[k, i, j, :]).
#pragma once
#include "common.cuh"

// 1: the hot kernel fetches the grid cell of a warp cooperatively through shared memory (grid_eval below)
#ifndef DABRY_COOP_FETCH
#define DABRY_COOP_FETCH 0
#endif

#ifndef DABRY_TRACE_CELL  // analysis hook of tests/hostsim (cell of every grid sample); nothing in the product
#define DABRY_TRACE_CELL(frame, i0, j0)
#endif

namespace dabry {

// FIELD_GRID_F32: the same pair records stored in fp32 (48 B per node, three 16-byte loads); arithmetic stays fp64
enum FieldKind : int32_t { FIELD_PROGRAM = 0, FIELD_GRID = 1, FIELD_GRID_F32 = 2 };

enum LeafKind : int32_t {
    LEAF_UNIFORM = 0,       // p: value(2)                                    flowfield.py:601-620
    LEAF_STATE_LINEAR = 1,  // p: gradient(4 row-major), origin(2), value_origin(2)   :730-752, :783-804
    LEAF_GYRE = 2,          // p: center(2), kx, ky, ampl                     :807-833
    LEAF_VORTEX = 3,        // p: center(2), circulation                      :623-649
    LEAF_RANKINE = 4,       // p: center(2), circulation, radius | table rows (cx, cy, gamma, radius)  :652-727
    LEAF_CHERTOVSKIH = 5,   //                                                :1089-1099
    LEAF_TRAP = 6,          // p[0]: rel_wid; table rows (ff_val, cx, cy, radius)   :1035-1086
    LEAF_RADIAL_GAUSS = 7,  // p: center(2), radius, sdev, v_max              :860-903
    LEAF_TWO_SECTORS = 8,   // p: v_w1, v_w2, x_switch                        :565-598
    LEAF_LINEAR_T = 9,      // table (3 rows): gradient_diff_time(4), gradient_init(4), origin(2) + value_origin(2)  :755-780
    LEAF_RADIAL_GAUSS_T = 10,  // table: nt rows (cx, cy, radius, sdev) then nt rows (v_max, -, -, -)   :906-973
    LEAF_BAND_GAUSS = 11,   // p: origin(2), unit vect(2), ampl, sdev; forward-difference Jacobian, dx 1e-6   :976-999
    LEAF_BAND = 12,         // p: origin(2), unit vect(2), w_value(2), width; central differences, eps 1e-6   :1002-1032
    LEAF_GYRE_MSEAS = 13,   // p: A, eps, T; central differences, eps 1e-6    :1102-1134
};

struct FieldLeaf {
    int32_t kind;
    int32_t nt;         // > 0: parameters are a time table of nt rows x 4 doubles (FlowField._index, :152-175)
    int32_t table_off;  // offset (in doubles) of the table inside FieldDesc::tables
    int32_t pad_;
    double coeff;       // signed product of the scalings on the path from the tree root
    double t_start, t_end;
    double p[8];
};

struct GridDesc {
    const double* data;  // pair records, 12 doubles per (k, i, j)
    int32_t nt, nx, ny;
    int32_t unsteady;    // 0: 2-D interpolation (flowfield.py:451-452, 466-468); nt == 1
    double lo[3];        // bounds[:, 0] for (t, x, y); lo[0] unused when steady
    double spacing[3];   // (hi - lo) / (n - 1), flowfield.py:198-199
    double inv_spacing[3];
};

struct FieldDesc {
    int32_t kind;
    int32_t n_leaves;
    const double* tables;
    GridDesc grid;
    // WrapperFF (flowfield.py:522-544)
    double time_origin, scale_time, scale_length, wbl[2], scaler_speed, scaler_dspeed;
    FieldLeaf leaves[DABRY_MAX_LEAVES];
};

struct Flow {  // w and J at one point
    double w0, w1, j00, j01, j10, j11;
};

struct D2v {
    double x, y;
};

// ---------------------------------------------------------------------------------------------------------
// FlowField._index (flowfield.py:152-175): lower frame and lerp coefficient of a time-tabulated parameter set
DABRY_HD void time_index(const FieldLeaf& L, double t, int& i, double& alpha) {
    const int nt = L.nt;
    i = 0;
    alpha = 0.;
    if (nt <= 1) return;
    const double t_min = dmin(L.t_start, L.t_end), t_max = dmax(L.t_start, L.t_end);
    if (t <= t_min) return;
    if (t > t_max) {
        i = nt - 2;
        alpha = 1.;
        return;
    }
    const double tau = (t - t_min) / (t_max - t_min);
    const double s = tau * (nt - 1);
    i = (int) s;
    alpha = s - (double) i;
    if (i == nt - 1) {
        i = nt - 2;
        alpha = 1.;
    }
}

// Value of the leaves whose Jacobian the reference takes by finite differences (BandGaussFF, BandFF, GyreMSEASFF).
// The differences amplify one ulp of the value 1e6 times, so products and sums are rounded one by one like numpy's.
DABRY_HD_NOINLINE D2v fd_leaf_value(const FieldLeaf& L, double t, double x0, double x1) {
    D2v o = {0., 0.};
    if (L.kind == LEAF_BAND_GAUSS || L.kind == LEAF_BAND) {
        const double d0 = sub_rn(x0, L.p[0]), d1 = sub_rn(x1, L.p[1]);
        const double dist = fabs(sub_rn(mul_rn(L.p[2], d1), mul_rn(L.p[3], d0)));  // |cross(vect, x - origin)|
        if (L.kind == LEAF_BAND_GAUSS) {
            const double q = dist / L.p[5];
            const double intensity = mul_rn(L.p[4], exp(mul_rn(-0.5, mul_rn(q, q))));
            o.x = mul_rn(L.p[2], intensity);
            o.y = mul_rn(L.p[3], intensity);
        } else {
            const double width = L.p[6], tw = width / 10;
            const double outer = add_rn(width, tw / 2) / 2, inner = sub_rn(width, tw / 2) / 2;
            if (dist >= outer) return o;
            double f = 1.;
            if (dist >= inner) f = sub_rn(1., sub_rn(dist, inner) / tw);
            o.x = dist >= inner ? mul_rn(f, L.p[4]) : L.p[4];
            o.y = dist >= inner ? mul_rn(f, L.p[5]) : L.p[5];
        }
    } else {  // LEAF_GYRE_MSEAS
        const double A = L.p[0], eps = L.p[1], T = L.p[2];
        const double a = mul_rn(eps, sin(mul_rn(mul_rn(2., M_PI), t) / T));
        const double b = sub_rn(1., mul_rn(2., a));
        const double f = add_rn(mul_rn(a, mul_rn(x0, x0)), mul_rn(b, x0));
        const double dfdx = add_rn(mul_rn(mul_rn(2., a), x0), b);
        const double pa = mul_rn(M_PI, A);
        o.x = mul_rn(mul_rn(-pa, sin(mul_rn(M_PI, f))), cos(mul_rn(M_PI, x1)));
        o.y = mul_rn(mul_rn(mul_rn(pa, cos(mul_rn(M_PI, f))), sin(mul_rn(M_PI, x1))), dfdx);
    }
    return o;
}

DABRY_HD void leaf_eval(const FieldLeaf& L, const double* __restrict__ tables, double t, double x0, double x1,
                        Flow& o) {
    o.w0 = o.w1 = o.j00 = o.j01 = o.j10 = o.j11 = 0.;
    switch (L.kind) {
    case LEAF_TWO_SECTORS:  // np.heaviside(., 0.): 0 at the switch itself; the Jacobian is zero as written (:586-588)
        o.w1 = L.p[0] * (L.p[2] - x0 > 0. ? 1. : 0.) + L.p[1] * (x0 - L.p[2] > 0. ? 1. : 0.);
        break;
    case LEAF_LINEAR_T: {
        const double* q = tables + L.table_off;  // gradient_diff_time, gradient_init, origin, value_origin
        const double m00 = q[0] * t + q[4], m01 = q[1] * t + q[5], m10 = q[2] * t + q[6], m11 = q[3] * t + q[7];
        const double d0 = x0 - q[8], d1 = x1 - q[9];
        o.w0 = (m00 * d0 + m01 * d1) + q[10];
        o.w1 = (m10 * d0 + m11 * d1) + q[11];
        o.j00 = m00;
        o.j01 = m01;
        o.j10 = m10;
        o.j11 = m11;
        break;
    }
    case LEAF_BAND_GAUSS: {  // forward differences, dx = 1e-6 (:995-999)
        const double dx = 1e-6, k = 1 / dx;
        const D2v v = fd_leaf_value(L, t, x0, x1);
        const D2v vx = fd_leaf_value(L, t, add_rn(x0, dx), x1), vy = fd_leaf_value(L, t, x0, add_rn(x1, dx));
        o.w0 = v.x;
        o.w1 = v.y;
        o.j00 = mul_rn(k, sub_rn(vx.x, v.x));
        o.j10 = mul_rn(k, sub_rn(vx.y, v.y));
        o.j01 = mul_rn(k, sub_rn(vy.x, v.x));
        o.j11 = mul_rn(k, sub_rn(vy.y, v.y));
        break;
    }
    case LEAF_BAND:
    case LEAF_GYRE_MSEAS: {  // central differences, eps = 1e-6 (:1026-1032, :1128-1134)
        const double eps = 1e-6, den = mul_rn(2., eps);
        const D2v v = fd_leaf_value(L, t, x0, x1);
        const D2v xp = fd_leaf_value(L, t, add_rn(x0, eps), x1), xm = fd_leaf_value(L, t, sub_rn(x0, eps), x1);
        const D2v yp = fd_leaf_value(L, t, x0, add_rn(x1, eps)), ym = fd_leaf_value(L, t, x0, sub_rn(x1, eps));
        o.w0 = v.x;
        o.w1 = v.y;
        o.j00 = sub_rn(xp.x, xm.x) / den;
        o.j10 = sub_rn(xp.y, xm.y) / den;
        o.j01 = sub_rn(yp.x, ym.x) / den;
        o.j11 = sub_rn(yp.y, ym.y) / den;
        break;
    }
    case LEAF_UNIFORM:
        o.w0 = L.p[0];
        o.w1 = L.p[1];
        break;
    case LEAF_STATE_LINEAR: {
        const double d0 = x0 - L.p[4], d1 = x1 - L.p[5];
        o.w0 = (L.p[0] * d0 + L.p[1] * d1) + L.p[6];
        o.w1 = (L.p[2] * d0 + L.p[3] * d1) + L.p[7];
        o.j00 = L.p[0];
        o.j01 = L.p[1];
        o.j10 = L.p[2];
        o.j11 = L.p[3];
        break;
    }
    case LEAF_GYRE: {
        const double kx = L.p[2], ky = L.p[3], ampl = L.p[4];
        const double a = kx * (x0 - L.p[0]), b = ky * (x1 - L.p[1]);
        double sa, ca, sb, cb;
        sincos(a, &sa, &ca);
        sincos(b, &sb, &cb);
        o.w0 = ampl * (-sa * cb);
        o.w1 = ampl * (ca * sb);
        o.j00 = ampl * (-kx * ca * cb);
        o.j01 = ampl * (ky * sa * sb);
        o.j10 = ampl * (-kx * sa * sb);
        o.j11 = ampl * (ky * ca * cb);
        break;
    }
    case LEAF_VORTEX: {
        const double dx = x0 - L.p[0], dy = x1 - L.p[1];
        const double r = sqrt(dx * dx + dy * dy);
        const double c = L.p[2] / (2 * M_PI * r);
        o.w0 = c * (-dy / r);
        o.w1 = c * (dx / r);
        const double r2 = r * r;
        const double g = L.p[2] / (2 * M_PI * (r2 * r2));
        o.j00 = g * (2 * dx * dy);
        o.j01 = g * (dy * dy - dx * dx);
        o.j10 = o.j01;
        o.j11 = g * (-2 * dx * dy);
        break;
    }
    case LEAF_RANKINE: {
        double cx, cy, gamma, radius;
        if (L.nt > 0) {
            int i;
            double al;
            time_index(L, t, i, al);
            const double* r0 = tables + L.table_off + 4 * i;
            const double* r1 = r0 + 4;
            cx = (1 - al) * r0[0] + al * r1[0];
            cy = (1 - al) * r0[1] + al * r1[1];
            gamma = (1 - al) * r0[2] + al * r1[2];
            radius = (1 - al) * r0[3] + al * r1[3];
        } else {
            cx = L.p[0];
            cy = L.p[1];
            gamma = L.p[2];
            radius = L.p[3];
        }
        const double a = x0 - cx, b = x1 - cy;
        const double r = sqrt(a * a + b * b);
        if (r < 1e-3 * radius) break;  // zero_ceil (flowfield.py:677, 695, 708)
        const double et0 = -b / r, et1 = a / r;
        const double f = gamma / (2 * M_PI);
        if (r <= radius) {
            const double s = f * r / (radius * radius);
            o.w0 = s * et0;
            o.w1 = s * et1;
            const double r3 = r * r * r;
            const double g = f / (radius * radius);
            // columns: a / r * e_theta + r * d e_theta/dx ,  b / r * e_theta + r * d e_theta/dy
            o.j00 = g * (a / r * et0 + r * (a * b / r3));
            o.j10 = g * (a / r * et1 + r * (b * b / r3));
            o.j01 = g * (b / r * et0 + r * (-(a * a) / r3));
            o.j11 = g * (b / r * et1 + r * (-(a * b) / r3));
        } else {
            const double s = f * 1. / r;
            o.w0 = s * et0;
            o.w1 = s * et1;
            const double r2 = r * r;
            const double g = f / (r2 * r2);
            o.j00 = g * (2 * a * b);
            o.j01 = g * (b * b - a * a);
            o.j10 = o.j01;
            o.j11 = g * (-2 * a * b);
        }
        break;
    }
    case LEAF_CHERTOVSKIH:
        o.w0 = x0 * (x1 + 3) / 4;
        o.w1 = -x0 * x0;
        o.j00 = (x1 + 3) / 4;
        o.j01 = x0 / 4;
        o.j10 = -2 * x0;
        o.j11 = 0.;
        break;
    case LEAF_TRAP: {
        int i;
        double al;
        time_index(L, t, i, al);
        const double* r0 = tables + L.table_off + 4 * i;
        const double* r1 = r0 + 4;
        const bool lo_only = al < 1e-3;  // flowfield.py:1064-1066
        const double ffv = (1 - al) * r0[0] + (lo_only ? 0. : al * r1[0]);
        const double cx = (1 - al) * r0[1] + (lo_only ? 0. : al * r1[1]);
        const double cy = (1 - al) * r0[2] + (lo_only ? 0. : al * r1[2]);
        const double radius = (1 - al) * r0[3] + (lo_only ? 0. : al * r1[3]);
        const double a = x0 - cx, b = x1 - cy;
        const double d2 = a * a + b * b;
        if (d2 < 1e-8 * (radius * radius)) break;
        const double r = sqrt(d2);
        const double e0 = a / r, e1 = b / r;
        const double wid = radius * L.p[0];
        const double s = 1. / (1. + exp(-4. / wid * (r - radius)));
        const double ds = (4. / wid) * s * (1. - s);
        o.w0 = -ffv * s * e0;
        o.w1 = -ffv * s * e1;
        // P = ((e0, e1), (-e1 / r, e0 / r));  J = -ff_val * P^T diag(ds, s) P
        const double q0 = -e1 / r, q1 = e0 / r;
        o.j00 = -ffv * (e0 * ds * e0 + q0 * s * q0);
        o.j01 = -ffv * (e0 * ds * e1 + q0 * s * q1);
        o.j10 = -ffv * (e1 * ds * e0 + q1 * s * q0);
        o.j11 = -ffv * (e1 * ds * e1 + q1 * s * q1);
        break;
    }
    case LEAF_RADIAL_GAUSS:
    case LEAF_RADIAL_GAUSS_T: {
        double cx = L.p[0], cy = L.p[1], radius = L.p[2], sdev = L.p[3], vmax = L.p[4];
        if (L.kind == LEAF_RADIAL_GAUSS_T) {  // parameters lerped between the frames of FlowField._index (:937-950)
            int i;
            double al;
            time_index(L, t, i, al);
            const double* r0 = tables + L.table_off + 4 * i;
            const double* r1 = r0 + 4;
            const double* v0 = tables + L.table_off + 4 * (L.nt + i);
            cx = (1 - al) * r0[0] + al * r1[0];
            cy = (1 - al) * r0[1] + al * r1[1];
            radius = (1 - al) * r0[2] + al * r1[2];
            sdev = (1 - al) * r0[3] + al * r1[3];
            vmax = (1 - al) * v0[0] + al * v0[4];
        }
        const double a = x0 - cx, b = x1 - cy;
        const double r = sqrt(a * a + b * b);
        if (r < 1e-3 * radius) break;
        const double e0 = a / r, e1 = b / r;
        const double lg = log(r / radius);
        const double amp = vmax * exp(-(lg * lg) / (2 * sdev * sdev));
        o.w0 = amp * e0;
        o.w1 = amp * e1;
        const double c = -lg * amp / (r * r * sdev * sdev);
        const double dv0 = c * a, dv1 = c * b;
        const double r3 = r * r * r;
        o.j00 = e0 * dv0 + amp * (b * b / r3);
        o.j01 = e0 * dv1 + amp * (-a * b / r3);
        o.j10 = e1 * dv0 + amp * (-a * b / r3);
        o.j11 = e1 * dv1 + amp * (a * a / r3);
        break;
    }
    default:
        break;
    }
}

// Node loads: on the device the six 16-byte halves of a pair record go through the read-only path.
struct D2 {
    double x, y;
};
DABRY_HD D2 load_pair(const double* p) {
#if defined(__CUDA_ARCH__)
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return D2{v.x, v.y};
#else
    return D2{p[0], p[1]};
#endif
}

struct F4 {
    float x, y, z, w;
};
DABRY_HD F4 load_quad(const float* p) {
#if defined(__CUDA_ARCH__)
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return F4{v.x, v.y, v.z, v.w};
#else
    return F4{p[0], p[1], p[2], p[3]};
#endif
}

// Utils.interpolate (misc.py:570-589) on values and grad_values at once.  Weights come from the UNCLIPPED
// floor; indices lo and lo + 1 are clipped independently to [0, n - 1]; no wrap in longitude.
//
// Frame records: record frame f = k + 1 holds (c[max(k, 0)], c[min(k + 1, nt - 1)]) for k = -1 .. nt - 1, so that the
// clipped pair (lo, hi) of ANY time position is one record and the time lerp needs no select: k < 0 -> both frame 0,
// k >= nt - 1 -> both frame nt - 1.  Positions use the host-computed reciprocal spacing (one rounding away from the
// reference's division; continuous in the result).
// JAC = false: only the two value components are sampled (boundary-following arcs and events need no Jacobian).
template <bool F32, bool JAC>
DABRY_HD void grid_eval(const GridDesc& G, double t, double x0, double x1, Flow& o) {
    double wt_lo = 1., wt_hi = 0.;
    int frame = 1;
    if (G.unsteady) {
        const double pos = (t - G.lo[0]) * G.inv_spacing[0];
        const double fl = floor(pos);
        wt_hi = pos - fl;
        wt_lo = 1. - wt_hi;
        frame = iclamp((int) fl, -1, G.nt - 1) + 1;
    }
    const double px = (x0 - G.lo[1]) * G.inv_spacing[1];
    const double py = (x1 - G.lo[2]) * G.inv_spacing[2];
    const double fx = floor(px), fy = floor(py);
    const double wx_hi = px - fx, wy_hi = py - fy;
    const double wx_lo = 1. - wx_hi, wy_lo = 1. - wy_hi;
    const int ix = (int) fx, iy = (int) fy;
    const int i0 = iclamp(ix, 0, G.nx - 1), i1 = iclamp(ix + 1, 0, G.nx - 1);
    const int j0 = iclamp(iy, 0, G.ny - 1), j1 = iclamp(iy + 1, 0, G.ny - 1);
    // record indices fit 32 bits (checked at upload); one 64-bit multiply per node for the byte offset
    DABRY_TRACE_CELL(frame, i0, j0);
    const unsigned row0 = ((unsigned) frame * (unsigned) G.nx + (unsigned) i0) * (unsigned) G.ny;
    const unsigned row1 = ((unsigned) frame * (unsigned) G.nx + (unsigned) i1) * (unsigned) G.ny;
    // corner weights (w_t * w_x) * w_y in the reference; here the time factor is applied per node last
    const double w00 = wx_lo * wy_lo, w01 = wx_lo * wy_hi, w10 = wx_hi * wy_lo, w11 = wx_hi * wy_hi;
    double acc[6] = {0., 0., 0., 0., 0., 0.};
    if (F32) {
        const float* base = reinterpret_cast<const float*>(G.data);
        const float* n00 = base + (size_t) (row0 + j0) * 12;
        const float* n01 = base + (size_t) (row0 + j1) * 12;
        const float* n10 = base + (size_t) (row1 + j0) * 12;
        const float* n11 = base + (size_t) (row1 + j1) * 12;
        // dtype F32: grid positions and weights are fp64 (a float position would quantise the weights to 1e-5 of a
        // cell), the 48 multiply-adds of the interpolation run in fp32 on the fp32 node data
        const float f00 = (float) w00, f01 = (float) w01, f10 = (float) w10, f11 = (float) w11;
        const float ft_lo = (float) wt_lo, ft_hi = (float) wt_hi;
#pragma unroll
        for (int c = 0; c < (JAC ? 3 : 1); ++c) {  // one float4 = two components x (frame k, frame k + 1)
            const F4 a = load_quad(n00 + 4 * c), b = load_quad(n01 + 4 * c);
            const F4 d = load_quad(n10 + 4 * c), e = load_quad(n11 + 4 * c);
            const float lo0 = ((f00 * a.x + f01 * b.x) + f10 * d.x) + f11 * e.x;
            const float hi0 = ((f00 * a.y + f01 * b.y) + f10 * d.y) + f11 * e.y;
            const float lo1 = ((f00 * a.z + f01 * b.z) + f10 * d.z) + f11 * e.z;
            const float hi1 = ((f00 * a.w + f01 * b.w) + f10 * d.w) + f11 * e.w;
            acc[2 * c] = (double) (ft_lo * lo0 + ft_hi * hi0);
            acc[2 * c + 1] = (double) (ft_lo * lo1 + ft_hi * hi1);
        }
    } else {
#if defined(__CUDA_ARCH__) && DABRY_COOP_FETCH
        if (JAC) {
            // Warp-cooperative cell fetch (north star: shared-memory staging of grid tiles for theta-sorted bundles).
            // The extremals of a warp are theta_0-neighbours: nearly always they sit in the same cell of the same pair of
            // frames, so the 24 sixteen-byte pieces of that cell are fetched ONCE per warp - one cp.async per lane into a
            // 384-byte tile of the warp in shared memory, all in flight at once and landing in no register - and every
            // lane then reads them as shared-memory broadcasts.  Lanes in another cell (a warp straddling a cell edge, a
            // lane that rejected a step and sits in the next frame) are served by further rounds of the same loop.
            __shared__ __align__(16) double2 coop_tile[4][24];  // arc kernels: at most 4 warps per block
            const unsigned active = __activemask();
            const int lane = threadIdx.x & 31;
            double2* tile = coop_tile[(threadIdx.x >> 5) & 3];
            const unsigned k00 = row0 + j0, k01 = row0 + j1, k10 = row1 + j0, k11 = row1 + j1;
            const int rank = __popc(active & ((1u << lane) - 1u)), n_active = __popc(active);
            unsigned pending = active;
            while (pending) {
                const int leader = __ffs(pending) - 1;
                const unsigned l00 = __shfl_sync(active, k00, leader), l01 = __shfl_sync(active, k01, leader);
                const unsigned l10 = __shfl_sync(active, k10, leader), l11 = __shfl_sync(active, k11, leader);
                const bool mine = ((pending >> lane) & 1u) && k00 == l00 && k11 == l11;
                for (int p = rank; p < 24; p += n_active) {
                    const unsigned node = p < 6 ? l00 : (p < 12 ? l01 : (p < 18 ? l10 : l11));
                    const double* src = G.data + (size_t) node * 12 + 2 * (p % 6);
                    const unsigned dst = (unsigned) __cvta_generic_to_shared(tile + p);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp(active);
                if (mine) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        const double2 a = tile[c], b = tile[6 + c], d = tile[12 + c], e = tile[18 + c];
                        const double lo = ((w00 * a.x + w01 * b.x) + w10 * d.x) + w11 * e.x;
                        const double hi = ((w00 * a.y + w01 * b.y) + w10 * d.y) + w11 * e.y;
                        acc[c] = wt_lo * lo + wt_hi * hi;
                    }
                }
                __syncwarp(active);
                pending &= ~__ballot_sync(active, mine);
            }
        } else
#endif
        {
        const double* n00 = G.data + (size_t) (row0 + j0) * 12;
        const double* n01 = G.data + (size_t) (row0 + j1) * 12;
        const double* n10 = G.data + (size_t) (row1 + j0) * 12;
        const double* n11 = G.data + (size_t) (row1 + j1) * 12;
        // Measured on B200 and dropped: staging the 24 pieces PER THREAD in shared memory with cp.async (3x
        // slower: the 48 KB per block shrink the L1 the integrator's spills live in), ordered inline-PTX loads (ptxas
        // re-schedules them), and prefetch.global.L1 of the cell ahead of the latitude trigonometry (+6 %).
#pragma unroll
        for (int c = 0; c < (JAC ? 6 : 2); ++c) {
            const D2 a = load_pair(n00 + 2 * c), b = load_pair(n01 + 2 * c);
            const D2 d = load_pair(n10 + 2 * c), e = load_pair(n11 + 2 * c);
            const double lo = ((w00 * a.x + w01 * b.x) + w10 * d.x) + w11 * e.x;
            const double hi = ((w00 * a.y + w01 * b.y) + w10 * d.y) + w11 * e.y;
            acc[c] = wt_lo * lo + wt_hi * hi;
        }
        }
    }
    o.w0 = acc[0];
    o.w1 = acc[1];
    o.j00 = acc[2];
    o.j01 = acc[3];
    o.j10 = acc[4];
    o.j11 = acc[5];
}

// WrapperFF.value / d_value (flowfield.py:538-544) around either family.
template <int FK, bool JAC = true>
DABRY_HD void field_eval(const FieldDesc& F, double t, double x0, double x1, Flow& o) {
    const double tt = F.time_origin + t * F.scale_time;
    const double xx0 = F.wbl[0] + x0 * F.scale_length;
    const double xx1 = F.wbl[1] + x1 * F.scale_length;
    if (FK == FIELD_GRID) {
        grid_eval<false, JAC>(F.grid, tt, xx0, xx1, o);
    } else if (FK == FIELD_GRID_F32) {
        grid_eval<true, JAC>(F.grid, tt, xx0, xx1, o);
    } else {
        Flow acc = {0., 0., 0., 0., 0., 0.};
        for (int l = 0; l < F.n_leaves; ++l) {
            Flow f;
            leaf_eval(F.leaves[l], F.tables, tt, xx0, xx1, f);
            const double c = F.leaves[l].coeff;
            if (l == 0) {
                acc.w0 = c * f.w0; acc.w1 = c * f.w1;
                acc.j00 = c * f.j00; acc.j01 = c * f.j01; acc.j10 = c * f.j10; acc.j11 = c * f.j11;
            } else {
                acc.w0 += c * f.w0; acc.w1 += c * f.w1;
                acc.j00 += c * f.j00; acc.j01 += c * f.j01; acc.j10 += c * f.j10; acc.j11 += c * f.j11;
            }
        }
        o = acc;
    }
    o.w0 *= F.scaler_speed;
    o.w1 *= F.scaler_speed;
    o.j00 *= F.scaler_dspeed;
    o.j01 *= F.scaler_dspeed;
    o.j10 *= F.scaler_dspeed;
    o.j11 *= F.scaler_dspeed;
}

}  // namespace dabry
