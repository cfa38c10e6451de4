// The solver loops of the reference (dabry/solver_ef.py) as a sequence of data-parallel phases over the site
// store.  Every phase is a functor `void operator()(int64_t i)` handed to `Backend::launch(n, functor)`:
// the product library instantiates Engine<CudaBackend> (dabry_b200.cu: one CUDA thread per i), tests/hostsim
// instantiates Engine<SerialBackend> with g++ to step through the same source on the CPU (test infrastructure).
//
//   SolverEFResampling.setup / step / solve     solver_ef.py:462-470, 582-608   -> setup(), step(), solve()
//   compute_new_sites                           solver_ef.py:630-665             -> ResamplePhase + order-preserving
//                                                                                   compaction + SpawnPhase
//   trim_distance                               solver_ef.py:594-600             -> TrimDistancePhase
//   SolverEFTrimming.setup / step / trim        solver_ef.py:778-841             -> step_trimming(), trim()
//   check_solutions                             solver_ef.py:667-685             -> finish()
#pragma once
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dabry_b200.h"
#include "site.cuh"

// RK steps per launch of the free-arc kernel, with re-packing of the running extremals in between; 0 = one launch.
// Measured on B200 (era5, 1.25 M extremals): 0 -> 25.9 ms, 20 -> 28.4, 10 -> 31.0, 5 -> 41.8: the re-packing costs
// more (lost L1 / theta_0 locality, carry traffic, launch tails) than the idle lanes it removes, so it is off.
#ifndef DABRY_FREE_CHUNK
#define DABRY_FREE_CHUNK 0
#endif
#ifndef DABRY_GRID_GROUPS  // > 0: fixed number of groups of time frames of the streamed grid upload (0: by frame count)
#define DABRY_GRID_GROUPS 0
#endif
#ifndef DABRY_INTEGRATE_MIN_BLOCKS
#define DABRY_INTEGRATE_MIN_BLOCKS 4
#endif

namespace dabry {

static_assert(sizeof(FieldLeaf) == sizeof(dabry_leaf), "leaf layout");
static_assert(sizeof(ObstacleDesc) == sizeof(dabry_obstacle), "obstacle layout");
static_assert(sizeof(PenaltyDesc) == sizeof(dabry_penalty), "penalty layout");
static_assert(DABRY_MAX_TARGET_EVENTS == 4, "abi");

struct Counters {  // device-resident
    unsigned long long rk_steps, ivp_calls, rk_steps_free, spare1;  // rk_steps_free: the free-arc kernel's share
};

DABRY_HD int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
DABRY_HD int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

DABRY_HD void counter_add(unsigned long long* c, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(c, v);
#else
    *c += v;
#endif
}

DABRY_HD void counter_max(unsigned long long* c, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicMax(c, v);
#else
    if (v > *c) *c = v;
#endif
}

// ---- phases -----------------------------------------------------------------------------------------------

struct FillI32Phase {
    static constexpr int kThreads = 256;
    int32_t* p;
    int32_t v;
    DABRY_HD void operator()(int64_t i) const { p[i] = v; }
};

struct FillF64Phase {
    static constexpr int kThreads = 256;
    double* p;
    double v;
    DABRY_HD void operator()(int64_t i) const { p[i] = v; }
};

// SolverEFResampling.setup (solver_ef.py:462-470): the ring of root extremals
struct InitRootsPhase {
    static constexpr int kThreads = 256;
    SiteStore S;
    const double* costates;  // [count][2] or null
    double x0, x1, t_init, dtheta;
    int64_t root_offset, root_count, n_roots;
    int64_t block, block_stride;  // local root j is ring root root_offset + (j / block) * block_stride + j % block
    int32_t max_depth, has_ghost;
    DABRY_HD int64_t ring_index(int64_t j) const { return root_offset + (j / block) * block_stride + j % block; }
    DABRY_HD void operator()(int64_t i) const {
        const int s = (int) i;
        site_reset(S, s);
        const int64_t g = ring_index(i);
        double p0, p1;
        if (costates) {
            p0 = costates[2 * i];
            p1 = costates[2 * i + 1];
        } else {
            // theta = linspace(0, 2 pi, n, endpoint=False)[g] = g * (2 pi / n); costate 1000 (cos, sin) (:465-467)
            const double th = mul_rn((double) g, dtheta);
            double sn, cs;
            sincos(th, &sn, &cs);
            p0 = 1000. * cs;
            p1 = 1000. * sn;
        }
        const int64_t o = S.at(s, 0);
        S.xp[o] = Sample{x0, x1, p0, p1};
        S.tt[o] = t_init;
        arrival_note(problem(), S, s, 0, x0, x1);  // a start inside the target disc arrives at cost 0
        S.dyadic[s] = g << max_depth;
        S.valid_first[s] = 0;
        S.valid_last[s] = 0;
        // ring neighbour at grid index 0 (:468-469); when the ring is sharded over several handles the last root of
        // every local block points at that block's ghost slot (a copy of the next rank's root, slots after the roots)
        int nb = s + 1;
        if (has_ghost) {
            if (i % block == block - 1) nb = (int) (root_count + i / block);
        } else if (i == root_count - 1) {
            nb = 0;
        }
        S.next_nb[o] = nb;
    }
};

struct GhostInitPhase {  // ghost k = the ring root after the last root of local block k, not yet integrated
    static constexpr int kThreads = 64;
    InitRootsPhase roots;
    DABRY_HD void operator()(int64_t k) const {
        const int64_t slot = roots.root_count + k;
        const int64_t g = (roots.ring_index(k * roots.block + roots.block - 1) + 1) % roots.n_roots;
        InitRootsPhase h = roots;
        h.costates = nullptr;
        h.has_ghost = 0;
        h.block = (int64_t) 1 << 60;  // identity mapping: ring index = root_offset + slot
        h.block_stride = 0;
        h.root_offset = g - slot;
        h.root_count = -1;
        h(slot);
        h.S.flags[slot] = SITE_GHOST;
        h.S.next_nb[h.S.at((int) slot, 0)] = -1;
    }
};

#ifndef DABRY_TRACE_FREE_ARC  // analysis hook of tests/hostsim (per-extremal attempt counts); nothing in the product
#define DABRY_TRACE_FREE_ARC(site, attempts)
#endif

// Small batches (the children a resampling round creates, the reference's own 30-root cases) are bound by the latency
// of ONE extremal's dependent chain, not by throughput: a few threads integrate up to 100 steps each while the rest of
// the GPU idles (config 5: 13 such launches of ~1 ms were 40 % of a trimming solve).  LEAN = false is the same source
// built for that case: one warp per block so that the batch spreads over all SMs, no register cap (255: every node
// load of a right-hand side in flight at once, no spill on the critical path).
#ifndef DABRY_WIDE_BATCH  // work items up to which the latency build is launched
#define DABRY_WIDE_BATCH 16384
#endif

template <int FK, int METHOD, bool CHUNKED, bool LEAN = true>
struct FreeArcPhase {  // the hot kernel: one thread per extremal, free arc with events, `budget` RK steps per launch
    static constexpr int kThreads = LEAN ? DABRY_FREE_ARC_THREADS : 32;
    static constexpr int kMinBlocks = LEAN ? DABRY_INTEGRATE_MIN_BLOCKS : 1;  // register cap = 65536 / (128 * kMinBlocks)
    SiteStore S;
    const int32_t* list;  // null: sites [base, base + n)
    int32_t base, index_t, budget;
    Counters* counters;
    int32_t* running;     // -1 / 0 per work item: the arc is parked and continues in the next chunk
    double t_park = INFINITY;  // CHUNKED: park the arc between two steps once its time reaches t_park (streamed grid)
    int32_t resume_only = 0;   // CHUNKED: only continue parked arcs (a later window over the same work items)
    DABRY_HD void operator()(int64_t i) const {
        const int s = list ? list[i] : base + (int) i;
        unsigned attempts = 0, calls = 0;
        // the throughput build of a gridded field runs the rolled step (site.cuh DABRY_FAST_STEP)
        constexpr bool kFast = LEAN && DABRY_FAST_STEP != 0 && FK != FIELD_PROGRAM;
        site_free_arc<FK, METHOD, CHUNKED, kFast>(problem(), S, s, index_t, attempts, calls, budget, t_park,
                                                  resume_only != 0);
        DABRY_TRACE_FREE_ARC(s, attempts);
        counter_add(&counters->rk_steps, attempts);
        counter_add(&counters->rk_steps_free, attempts);
        counter_add(&counters->ivp_calls, calls);
        if (CHUNKED) running[i] = (S.flags[s] & SITE_RUNNING) ? 0 : -1;
    }
};

// Tiny batches (the reference's own 30-root cases: a resampling round integrates a few dozen children) are bound by the
// latency of one warp: 32 extremals in different branches of the field, the event handling and the step control
// serialise on it (ncu on `movor`: 12.7 of 32 lanes active per instruction).  Spread<Phase> gives every work item a warp
// of its own (lane 0 works, 31 lanes idle): the GPU has 4 700 warp slots, a round of a small case needs tens.
#ifndef DABRY_SPREAD_BATCH  // work items up to which an arc kernel runs one extremal per warp
#define DABRY_SPREAD_BATCH 2048
#endif
#ifndef DABRY_SPREAD_THREADS  // warps of a block of the one-extremal-per-warp launch share the instruction cache of an SM
#define DABRY_SPREAD_THREADS 32
#endif
template <class Phase>
struct Spread {
    static constexpr int kThreads = DABRY_SPREAD_THREADS;
    static constexpr int kMinBlocks = 1;
    Phase inner;
    DABRY_HD void operator()(int64_t i) const {
        if ((i & 31) == 0) inner(i >> 5);
    }
};

struct CapturedFlagPhase {  // which work items need the obstacle kernel
    static constexpr int kThreads = 256;
    SiteStore S;
    const int32_t* list;
    int32_t base;
    int32_t* captured;
    DABRY_HD void operator()(int64_t i) const {
        const int s = list ? list[i] : base + (int) i;
        captured[i] = (S.arc_from[s] >= 0 && S.obstacle[s] >= 0) ? 0 : -1;
    }
};

#ifndef DABRY_OBSTACLE_MIN_BLOCKS
#define DABRY_OBSTACLE_MIN_BLOCKS 4  // measured 2 / 3 / 4 / 5 blocks per SM: 2.95 / 2.93 / 2.64 / 2.60 ms (config 4), 8.59 / 8.56 / 8.22 / 8.46 ms (config 3)
#endif
template <int FK, int METHOD, bool LEAN = true>
struct ObstacleArcPhase {  // entry side, feasibility and boundary following for the captured extremals
    static constexpr int kThreads = LEAN ? 128 : 32;
    static constexpr int kMinBlocks = LEAN ? DABRY_OBSTACLE_MIN_BLOCKS : 1;
    SiteStore S;
    const int32_t* list;      // the integrate work list (null: base + item)
    const int32_t* captured;  // dense list of captured work items
    int32_t base, index_t;
    Counters* counters;
    DABRY_HD void operator()(int64_t i) const {
        const int item = captured[i];
        const int s = list ? list[item] : base + item;
        unsigned attempts = 0, calls = 0;
        site_obstacle_arcs<FK, METHOD>(problem(), S, s, index_t, attempts, calls);
        counter_add(&counters->rk_steps, attempts);
        counter_add(&counters->ivp_calls, calls);
    }
};

struct ConnectPhase {  // connect_to_parents (solver_ef.py:690-699)
    static constexpr int kThreads = 256;
    SiteStore S;
    const int32_t* list;
    int32_t base;
    DABRY_HD void operator()(int64_t i) const { site_connect(S, list ? list[i] : base + (int) i); }
};

struct ResamplePhase {  // compute_new_sites, one site of the loop at solver_ef.py:634-661
    static constexpr int kThreads = 256;
    SolverParams C;
    SiteStore S;
    int32_t index_t_hi;
    int32_t* request;
    DABRY_HD void operator()(int64_t i) const {
        const int s = (int) i;
        request[s] = (S.flags[s] & SITE_GHOST) ? -1 : site_resample(problem(), C, S, s, index_t_hi);
    }
};

struct SpawnPhase {  // site_from_parents + dict insertion order (solver_ef.py:651, 663-664)
    static constexpr int kThreads = 256;
    SolverParams C;
    SiteStore S;
    const int32_t* request;
    const int32_t* offset;  // exclusive prefix sum of (request >= 0)
    int32_t n_old, subframe;
    DABRY_HD void operator()(int64_t i) const {
        const int s = (int) i;
        if (request[s] >= 0) site_spawn(C, S, s, n_old + offset[s], request[s], subframe);
    }
};

struct TrimDistancePhase {
    static constexpr int kThreads = 256;
    SolverParams C;
    SiteStore S;
    DABRY_HD void operator()(int64_t i) const {
        if (!(S.flags[i] & SITE_GHOST)) site_trim_distance(problem(), C, S, (int) i);
    }
};

struct ArrivalPhase {
    static constexpr int kThreads = 256;
    SiteStore S;
    int32_t* arr_index;
    double* arr_cost;
    int32_t rescan;  // 1: recompute from the trajectories (site_arrival) instead of trusting the running records
    DABRY_HD void operator()(int64_t i) const {
        double c = 0.;
        int k = -1;
        if (!(S.flags[i] & SITE_GHOST)) {
            if (rescan) {
                k = site_arrival(problem(), S, (int) i, c);
            } else {
                k = S.arr_index[i];  // kept up to date by every writer of a sample (arrival_note)
                c = S.arr_cost[i];
            }
        }
        arr_index[i] = k;
        arr_cost[i] = k >= 0 ? c : INFINITY;
    }
};

struct SolutionFlagPhase {  // solution_sites (solver_ef.py:676-685): arrived, and not costlier than `bound`
    static constexpr int kThreads = 256;
    const int32_t* arr_index;
    const double* arr_cost;
    double bound;
    int32_t* flag;
    DABRY_HD void operator()(int64_t i) const { flag[i] = (arr_index[i] >= 0 && !(arr_cost[i] > bound)) ? 0 : -1; }
};

struct ValidFlagPhase {  // membership in sites_valid[subframe] (solver_ef.py:803)
    static constexpr int kThreads = 256;
    SiteStore S;
    int32_t subframe;
    int32_t* flag;
    DABRY_HD void operator()(int64_t i) const {
        flag[i] = (!(S.flags[i] & SITE_GHOST) && S.valid_first[i] <= subframe && subframe <= S.valid_last[i]) ? 0 : -1;
    }
};

struct RasterPhase {  // cost_map_triangle (solver_ef.py:860-889), one (site, grid index) cell per i
    static constexpr int kThreads = 128;
    SolverParams C;
    SiteStore S;
    double* map;
    int32_t n_sites, k_lo, k_hi;          // grid indices [k_lo, k_hi]
    int32_t subframe, valid_a, valid_b;   // site set: created at `subframe` or valid in some sub-frame of [a, b)
    int32_t all_sites;
    DABRY_HD void operator()(int64_t i) const {
        const int s = (int) (i % n_sites), k = k_lo + (int) (i / n_sites);
        if (S.flags[s] & SITE_GHOST) return;
        if (!all_sites) {
            const bool member = S.created_subframe[s] == subframe ||
                                (valid_a < valid_b && S.valid_first[s] <= valid_b - 1 && S.valid_last[s] >= valid_a);
            if (!member) return;
        }
        if (k < S.index_t_init[s] || k >= S.check_next[s] || k > k_hi) return;
        site_rasterize(C, S, map, s, k);
    }
};

struct TrimPhase {  // SolverEFTrimming.trim, loop at solver_ef.py:831-841
    static constexpr int kThreads = 256;
    SolverParams C;
    SiteStore S;
    const double* map;
    int32_t subframe, index;
    DABRY_HD void operator()(int64_t i) const {
        const int s = (int) i;
        if (S.flags[s] & SITE_GHOST) return;
        const bool considered = (S.valid_first[s] <= subframe && subframe <= S.valid_last[s]) ||
                                S.created_subframe[s] == subframe;
        if (!considered || S.closure[s] != CLOSURE_NONE) return;
        if (site_trim(C, S, map, s, index)) {
            if (S.valid_first[s] > S.valid_last[s]) S.valid_first[s] = subframe + 1;
            S.valid_last[s] = subframe + 1;
        }
    }
};

template <int FK>
struct EvalRhsPhase {
    static constexpr int kThreads = 128;
    const double* t;
    const double* y;
    double* dy;
    DABRY_HD void operator()(int64_t i) const { rhs_aug<FK>(problem(), t[i], y + 4 * i, dy + 4 * i); }
};

template <int FK>
struct EvalFlowPhase {
    static constexpr int kThreads = 128;
    const double* t;
    const double* x;
    double* w;
    double* jac;
    DABRY_HD void operator()(int64_t i) const {
        Flow f;
        field_eval<FK>(problem().field, t[i], x[2 * i], x[2 * i + 1], f);
        w[2 * i] = f.w0;
        w[2 * i + 1] = f.w1;
        jac[4 * i] = f.j00;
        jac[4 * i + 1] = f.j01;
        jac[4 * i + 2] = f.j10;
        jac[4 * i + 3] = f.j11;
    }
};

template <int FK>
struct TimeBoundPhase {  // auto_time_upper_bound: work item 0 = orthodromic flight, 1 = heading-to-target flight
    static constexpr int kThreads = 32;
    double x0, x1, t_start, t_end;
    double* out;
    Counters* counters;
    DABRY_HD void operator()(int64_t i) const {
        unsigned attempts = 0;
        out[i] = feedback_flight_time<FK>(problem(), (int) i, x0, x1, t_start, t_end, attempts);
        counter_add(&counters->spare1, attempts);
    }
};

// DiscreteFF.compute_derivatives (flowfield.py:474-504) fused with the repack into pair records (fields.cuh).
// One i per node (k, ix, iy) of the packed grid.
struct PackGridPhase {
    static constexpr int kThreads = 256;
    const double* values;  // (nt, nx, ny, 2)
    const double* grad;    // (nt, nx, ny, 2, 2) or null
    double* out;           // pair records
    int32_t nt, nx, ny;
    double two_dx, two_dy;
    double scaler_speed;             // WrapperFF._scaler_speed
    int32_t f32;                     // store the records in fp32 (dtype F32)
    unsigned long long* max_norm;    // bits of max |scaler * w| over the nodes (non-negative doubles order like integers)
    int64_t first_record;            // work item 0 is this record
    DABRY_HD void node(int k, int ix, int iy, double* c) const {
        const int64_t n = ((int64_t) k * nx + ix) * ny + iy;
        c[0] = values[2 * n];
        c[1] = values[2 * n + 1];
        if (grad) {
            c[2] = grad[4 * n];
            c[3] = grad[4 * n + 1];
            c[4] = grad[4 * n + 2];
            c[5] = grad[4 * n + 3];
            return;
        }
        // central differences at the nearest interior node: borders copy the adjacent interior line, corners the
        // diagonal interior node (flowfield.py:487-504)
        const int ie = iclamp(ix, 1, nx - 2), je = iclamp(iy, 1, ny - 2);
        const int64_t xp = ((int64_t) k * nx + ie + 1) * ny + je, xm = ((int64_t) k * nx + ie - 1) * ny + je;
        const int64_t yp = ((int64_t) k * nx + ie) * ny + je + 1, ym = ((int64_t) k * nx + ie) * ny + je - 1;
        c[2] = (values[2 * xp] - values[2 * xm]) / two_dx;
        c[3] = (values[2 * yp] - values[2 * ym]) / two_dy;
        c[4] = (values[2 * xp + 1] - values[2 * xm + 1]) / two_dx;
        c[5] = (values[2 * yp + 1] - values[2 * ym + 1]) / two_dy;
    }
    DABRY_HD void operator()(int64_t item) const {
        const int64_t i = item + first_record;  // a streamed upload packs the record frames group by group
        const int iy = (int) (i % ny);
        const int ix = (int) ((i / ny) % nx);
        const int k = (int) (i / ((int64_t) ny * nx)) - 1;  // record frame f = k + 1, k = -1 .. nt - 1
        double a[6], b[6];
        node(imax(k, 0), ix, iy, a);
        node(imin(k + 1, nt - 1), ix, iy, b);
        if (f32) {
            float* o = reinterpret_cast<float*>(out) + 12 * i;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                o[2 * c] = (float) a[c];
                o[2 * c + 1] = (float) b[c];
            }
        } else {
            double* o = out + 12 * i;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                o[2 * c] = a[c];
                o[2 * c + 1] = b[c];
            }
        }
        if (k >= 0) {
            // np.max(np.linalg.norm(scaler * values, axis=-1)) (solver_ef.py:459, flowfield.py:547-548), same roundings
            const double su = mul_rn(scaler_speed, a[0]), sv = mul_rn(scaler_speed, a[1]);
            const double norm = sqrt(add_rn(mul_rn(su, su), mul_rn(sv, sv)));
            unsigned long long key;
            memcpy(&key, &norm, sizeof(key));
            if (norm == norm && key > *max_norm) counter_max(max_norm, key);
        }
    }
};

struct GatherTrajPhase {  // time-major store -> per-site padded rows for the host (dabry_get_trajs)
    static constexpr int kThreads = 256;
    SiteStore S;
    int64_t first;
    int32_t n_time;
    double* times;
    double* states;
    double* costates;
    double* controls;
    double* cost;
    DABRY_HD void operator()(int64_t i) const {
        const int64_t r = i / n_time;
        const int k = (int) (i % n_time);
        const int s = (int) (first + r);
        const int k0 = S.index_t_init[s], n = S.n_samples[s];
        const bool ok = k >= k0 && k < k0 + n;
        const int64_t o = S.at(s, k);
        const double nan = quiet_nan();
        if (times) times[i] = ok ? S.tt[o] : nan;
        if (states) {
            states[2 * i] = ok ? S.xp[o].x0 : nan;
            states[2 * i + 1] = ok ? S.xp[o].x1 : nan;
        }
        if (costates) {
            costates[2 * i] = ok ? S.xp[o].p0 : nan;
            costates[2 * i + 1] = ok ? S.xp[o].p1 : nan;
        }
        if (controls) {
            controls[2 * i] = ok ? S.ctrl[o].u0 : nan;
            controls[2 * i + 1] = ok ? S.ctrl[o].u1 : nan;
        }
        if (cost) cost[i] = ok ? site_cost_at(S, s, k) : nan;
    }
};

struct GatherNbPhase {
    static constexpr int kThreads = 256;
    SiteStore S;
    int64_t first;
    int32_t n_time;
    int32_t* out;
    DABRY_HD void operator()(int64_t i) const {
        out[i] = S.next_nb[S.at((int) (first + i / n_time), (int) (i % n_time))];
    }
};

struct GatherRecordPhase {
    static constexpr int kThreads = 256;
    SiteStore S;
    int64_t first;
    dabry_site_record* out;
    DABRY_HD void operator()(int64_t i) const {
        const int s = (int) (first + i);
        dabry_site_record r;
        r.dyadic = S.dyadic[s];
        r.index_t_init = S.index_t_init[s];
        r.n_samples = S.n_samples[s];
        r.closure = S.closure[s];
        r.neuter = S.neuter[s];
        r.index_neutered = S.index_neutered[s];
        r.obstacle = S.obstacle[s];
        r.index_t_obs = S.index_t_obs[s];
        r.obs_trigo = S.trigo[s];
        r.index_t_check_next = S.check_next[s];
        r.depth = S.depth[s];
        r.parent_prev = S.parent_prev[s];
        r.parent_next = S.parent_next[s];
        r.n_target_events = S.n_target_events[s];
        r.ivp_status = S.ivp_status[s];
        r.flags = S.flags[s];
        r.pad_ = 0;
        r.t_enter_obs = S.t_enter_obs[s];
        r.cost0 = S.cost0[s];
        for (int e = 0; e < DABRY_MAX_TARGET_EVENTS; ++e)
            r.t_target_events[e] = S.t_target_events[(int64_t) s * DABRY_MAX_TARGET_EVENTS + e];
        out[i] = r;
    }
};

// multi-GPU ring boundary: one site <-> a packed buffer of n_time * 7 + 16 doubles
#define DABRY_BOUNDARY_HEADER 16
struct BoundaryPhase {  // one work item per (packet, time index)
    static constexpr int kThreads = 128;
    SiteStore S;
    int32_t do_import, n_time;
    int64_t root_count, block, packet;  // export: first root of local block k; import: ghost slot root_count + k
    double* all;
    // import: ghost k is packet (k + shift) % n_packets of the buffer.  The in-library ring exchange receives the next
    // rank's packets in that rank's block order: block k of rank r is followed by block k of rank r + 1, except on the
    // last rank, where it is followed by block k + 1 of rank 0 (shift 1)
    int64_t shift = 0, n_packets = 1;
    DABRY_HD void operator()(int64_t i) const {
        const int64_t pk = i / n_time;
        const int k = (int) (i % n_time);
        const int site = (int) (do_import ? root_count + pk : pk * block);
        double* buf = all + (do_import ? (pk + shift) % n_packets : pk) * packet;
        const int64_t o = S.at(site, k);
        double* row = buf + DABRY_BOUNDARY_HEADER + 7 * k;
        if (!do_import) {
            row[0] = S.xp[o].x0; row[1] = S.xp[o].x1; row[2] = S.xp[o].p0; row[3] = S.xp[o].p1;
            row[4] = S.ctrl[o].u0; row[5] = S.ctrl[o].u1; row[6] = S.tt[o];
            if (k == 0) {
                buf[0] = (double) S.index_t_init[site]; buf[1] = (double) S.n_samples[site];
                buf[2] = (double) S.closure[site]; buf[3] = (double) S.obstacle[site];
                buf[4] = (double) S.index_t_obs[site]; buf[5] = (double) S.trigo[site];
                buf[6] = (double) S.depth[site]; buf[7] = S.cost0[site];
                const int64_t dy = S.dyadic[site];  // raw bits: a dyadic index may exceed 2^53
                memcpy(&buf[8], &dy, sizeof(dy));
            }
        } else {
            S.xp[o] = Sample{row[0], row[1], row[2], row[3]};
            S.ctrl[o] = Control{row[4], row[5]};
            S.tt[o] = row[6];
            if (k == 0) {
                site_reset(S, site);
                S.index_t_init[site] = (int) buf[0]; S.n_samples[site] = (int) buf[1];
                S.closure[site] = (int) buf[2]; S.obstacle[site] = (int) buf[3];
                S.index_t_obs[site] = (int) buf[4]; S.trigo[site] = (int) buf[5];
                S.depth[site] = (int) buf[6]; S.cost0[site] = buf[7];
                int64_t dy;
                memcpy(&dy, &buf[8], sizeof(dy));
                S.dyadic[site] = dy;
                S.flags[site] = SITE_GHOST;
                S.check_next[site] = 0;
            }
        }
    }
};

// ---- engine -------------------------------------------------------------------------------------------------

template <class B>
class Engine {
public:
    B be;
    dabry_config cfg;
    ProblemDev P;  // host copy; bound to the device's constant memory by bind()
    SolverParams C;
    SiteStore S;
    std::string err;
    dabry_stats stats;

    explicit Engine(const dabry_config& c) : cfg(c) {
        memset(&P, 0, sizeof(P));
        memset(&C, 0, sizeof(C));
        memset(&S, 0, sizeof(S));
        memset(&stats, 0, sizeof(stats));
        if (const char* env = getenv("DABRY_FREE_CHUNK")) free_chunk = atoi(env);
        if (const char* env = getenv("DABRY_B200_WIDE_BATCH")) wide_batch = atoll(env);
        if (const char* env = getenv("DABRY_B200_SPREAD_BATCH")) spread_batch = atoll(env);
        if (spread_batch > wide_batch) spread_batch = wide_batch;  // WIDE_BATCH=0 pins every batch to the lean build
        if (const char* env = getenv("DABRY_B200_VERIFY_ARRIVALS")) verify_arrivals = atoi(env) != 0;
        if (const char* env = getenv("DABRY_B200_STREAM_UPLOAD")) stream_upload = atoi(env);
        times_h.assign(c.times, c.times + c.n_time);
        cfg.times = nullptr;
        P.coords = c.coords;
        P.srf = c.srf;
        P.x_target[0] = c.x_target[0];
        P.x_target[1] = c.x_target[1];
        P.target_radius_sq = c.target_radius * c.target_radius;  // ** 2 (solver_ef.py:312)
        P.rtol = c.rtol;
        P.atol = c.atol;
        P.max_step = c.max_step;
        P.field.kind = -1;
        P.penalty.kind = PEN_NONE;
        C.max_depth = c.max_depth;
        C.mode_origin = c.mode_origin;
        C.follow_obstacles = 1;
        C.n_roots = c.n_roots;
        C.cm_nx = c.cost_map_nx;
        C.cm_ny = c.cost_map_ny;
        C.max_dist_sq = c.max_dist * c.max_dist;
        C.ff_max_norm = c.ff_max_norm;
        for (int a = 0; a < 2; ++a) {
            C.bl[a] = c.bl[a];
            C.tr[a] = c.tr[a];
            C.cm_inv[a] = (double) ((a == 0 ? c.cost_map_nx : c.cost_map_ny) - 1) / (c.tr[a] - c.bl[a]);
        }
    }

    ~Engine() {
        drop_pending_grid();
        free_store(S);
        be.free(d_times);
        be.free(d_tables);
        be.free(d_pen_data);
        be.free(d_pen_axes);
        be.free(d_grid);
        be.free(d_request);
        be.free(d_offset);
        be.free(d_list);
        be.free(d_captured);
        be.free(d_cap_offset);
        be.free(d_cap_list);
        be.free(d_run_a);
        be.free(d_run_b);
        be.free(d_map);
        be.free(d_arr_index);
        be.free(d_arr_cost);
        be.free(d_counters);
        be.free(d_send);
        be.free(d_recv);
        be.free(d_scratch);
    }

    int init() {
        if (!be.ok()) return fail(DABRY_ERR_NO_DEVICE, be.error());
        d_times = (double*) be.alloc(sizeof(double) * cfg.n_time);
        be.upload(d_times, times_h.data(), sizeof(double) * cfg.n_time);
        d_counters = (Counters*) be.alloc(sizeof(Counters));
        be.zero(d_counters, sizeof(Counters));
        d_map = (double*) be.alloc(sizeof(double) * C.cm_nx * C.cm_ny);
        return check();
    }

    // ---- problem ----
    // shared: every rank of the attached communicator holds the same `values`; each uploads 1 / world of them and the
    // rest arrives from the peers over NVLink (all-gather in place) instead of world x the same bytes over PCIe
    int set_flow_grid(const double* values, const double* grad, int nt, int nx, int ny, int unsteady,
                      const double* lo, const double* hi, const dabry_wrapper* w, bool shared = false) {
        if (!values || nt < 1 || nx < 3 || ny < 3 || (!unsteady && nt != 1))
            return fail(DABRY_ERR_INVALID, "flow grid: need values, nt >= 1, nx >= 3, ny >= 3 (nt == 1 if steady)");
        drop_pending_grid();
        be.tic();
        const int64_t nodes = (int64_t) nt * nx * ny;
        const int64_t records = (int64_t) (nt + 1) * nx * ny;
        if (records >= (int64_t) 1 << 31) return fail(DABRY_ERR_CAPACITY, "flow grid: more than 2^31 records");
        const int world = shared ? comm_world() : 1;
        // all-gather needs equal shares: the staging buffer is padded to world x share doubles
        const int64_t share = (2 * nodes + world - 1) / world;
        double* d_values = (double*) be.alloc(sizeof(double) * (world > 1 ? share * world : 2 * nodes));
        double* d_grad = grad ? (double*) be.alloc(sizeof(double) * 4 * nodes) : nullptr;
        be.free(d_grid);
        d_grid = (double*) be.alloc((cfg.dtype == DABRY_F32 ? sizeof(float) : sizeof(double)) * 12 * records);
        if (!d_values || !d_grid || (grad && !d_grad)) return fail(DABRY_ERR_CAPACITY, "flow grid: out of device memory");
        // A page-locked `values` array of an unsteady field is uploaded in the background, a group of frames at a time
        // on the copy stream; the record frames are packed group by group when the first solve needs them and the free
        // arcs of that solve are parked at the time the next group starts (integrate()).  The caller keeps `values`
        // alive and unchanged until that solve (or any other call that needs the whole grid) returns.
        // (only grids whose copy takes longer than the extra launches of the windows cost: 64 MB is 1.3 ms of PCIe)
        const bool big = stream_upload > 1 || sizeof(double) * 2 * (size_t) nodes >= ((size_t) 64 << 20);
        const bool streamed = world == 1 && stream_upload && big && unsteady && !grad && nt >= 4 &&
                              (!w || w->scale_time > 0.) && be.host_is_pinned(values);
        if (world > 1) {
            const int64_t first = share * comm_rank();
            const int64_t mine = imax64(0, imin64(share, 2 * nodes - first));
            if (mine > 0) be.upload_async(d_values + first, values + first, sizeof(double) * mine);
            be.span_begin();
            be.comm_allgather(comm, d_values + first, d_values, sizeof(double) * share);
            be.span_end();
        } else if (!streamed) {
            be.upload(d_values, values, sizeof(double) * 2 * nodes);
        }
        if (grad) be.upload(d_grad, grad, sizeof(double) * 4 * nodes);
        GridDesc& G = P.field.grid;
        G.data = d_grid;
        G.nt = nt;
        G.nx = nx;
        G.ny = ny;
        G.unsteady = unsteady;
        const int n_ax[3] = {nt, nx, ny};
        for (int a = 0; a < 3; ++a) {
            G.lo[a] = lo[a];
            // (bounds[:, 1] - bounds[:, 0]) / (shape - 1)  (flowfield.py:198-199)
            G.spacing[a] = (a == 0 && !unsteady) ? 1. : (hi[a] - lo[a]) / (double) (n_ax[a] - 1);
            G.inv_spacing[a] = 1. / G.spacing[a];
        }
        unsigned long long* d_key = (unsigned long long*) be.alloc(sizeof(unsigned long long));
        be.zero(d_key, sizeof(unsigned long long));
        const int f32 = cfg.dtype == DABRY_F32 ? 1 : 0;
        PackGridPhase ph{d_values, d_grad, d_grid, nt, nx, ny, 2 * G.spacing[1], 2 * G.spacing[2],
                         w ? w->scaler_speed : 1., f32, d_key, 0};
        P.field.kind = f32 ? FIELD_GRID_F32 : FIELD_GRID;
        set_wrapper(w);
        if (streamed) {
            // about eight time frames per group: every group costs a window of the free arcs (a relaunch of every arc and
            // its tail, ~1 ms), the first group is waited for (0.36 ms per frame at the PCIe rate).  Measured with 3 / 4 / 5 /
            // 6 groups: 34.9 / 35.9 / 37.1 / 39.3 ms per end-to-end solve for config 4 (24 frames), 37.3 / 36.5 / 36.1 / 36.1
            // for config 3 (48 frames) - profiles/r02_ab_e2e_groups.txt.  DABRY_GRID_GROUPS > 0 fixes the number.
            n_groups = DABRY_GRID_GROUPS > 0 ? DABRY_GRID_GROUPS : imax(2, imin(kMaxGridGroups, (nt + 4) / 8));
            if (n_groups > nt / 2) n_groups = imax(1, nt / 2);
            const int per = (nt + n_groups - 1) / n_groups;
            for (int g = 0; g < n_groups; ++g) group_frames[g] = imin((g + 1) * per, nt);
            pending_pack = ph;
            pending_values = d_values;
            pending_host = values;
            pending_key = d_key;
            groups_packed = 0;
            groups_issued = 0;
            frames_packed = 0;
            grid_pending = true;
            // Only the first group is put on the copy engine now: host-to-device copies of all streams share one DMA
            // queue, and the small uploads of the solve (problem constants, root costates) must not wait behind the
            // whole grid.  Group g + 1 is issued when group g is packed, right before the window that hides it.
            issue_group();
            stats.ms_upload += be.toc();  // issue time only: the copies run behind the caller's back
            return check();
        }
        be.launch(records, ph);
        unsigned long long key = 0;
        be.download(&key, d_key, sizeof(key));
        be.free(d_key);
        memcpy(&flow_max_norm, &key, sizeof(key));
        if (cfg.ff_max_norm < 0.) C.ff_max_norm = flow_max_norm;  // "auto": the device-computed value
        be.sync();
        be.free(d_values);
        be.free(d_grad);
        stats.ms_upload += be.toc();
        return check();
    }

    // ---- streamed grid upload ----
    static constexpr int kMaxGridGroups = 8;  // = the copy-event slots of the backend
    void drop_pending_grid() {  // a handle destroyed or re-loaded while its grid is still arriving
        if (!grid_pending) return;
        be.sync_copies();
        be.free(pending_key);
        be.free(pending_values);
        pending_key = nullptr;
        pending_values = nullptr;
        grid_pending = false;
    }
    // pack the record frames whose two value frames have arrived with group `upto` (record frame f = k + 1 holds value
    // frames max(k, 0) and min(k + 1, nt - 1)); after the last group: max |w|, staging buffer released
    void issue_group() {  // background copy of the next group of value frames
        if (groups_issued >= n_groups) return;
        const GridDesc& G = P.field.grid;
        const int g = groups_issued++;
        const int f0 = g == 0 ? 0 : group_frames[g - 1], f1 = group_frames[g];
        const size_t frame_doubles = (size_t) 2 * G.nx * G.ny;
        if (f1 > f0)
            be.upload_chunk_async(pending_values + f0 * frame_doubles, pending_host + f0 * frame_doubles,
                                  sizeof(double) * frame_doubles * (size_t) (f1 - f0), g);
    }
    void pack_groups(int upto) {
        const GridDesc& G = P.field.grid;
        for (; groups_packed <= upto && groups_packed < n_groups; ++groups_packed) {
            const int have = group_frames[groups_packed];  // value frames 0 .. have - 1 are on the device after this group
            while (groups_issued <= groups_packed) issue_group();
            be.wait_chunk(groups_packed);
            issue_group();  // the next group travels while this one is packed and used
            const int hi = have >= G.nt ? G.nt + 1 : have;  // record frames [frames_packed, hi) are complete
            if (hi > frames_packed) {
                PackGridPhase ph = pending_pack;
                ph.first_record = (int64_t) frames_packed * G.nx * G.ny;
                be.launch((int64_t) (hi - frames_packed) * G.nx * G.ny, ph);
                frames_packed = hi;
            }
        }
        if (groups_packed >= n_groups && grid_pending) {
            unsigned long long key = 0;
            be.download(&key, pending_key, sizeof(key));
            memcpy(&flow_max_norm, &key, sizeof(key));
            if (cfg.ff_max_norm < 0.) C.ff_max_norm = flow_max_norm;
            be.free(pending_key);
            be.free(pending_values);
            pending_key = nullptr;
            pending_values = nullptr;
            grid_pending = false;
        }
    }
    void finish_grid() {
        if (grid_pending) pack_groups(n_groups - 1);
    }
    // solver time before which every right-hand side of a step only touches value frames 0 .. have - 1
    double park_time(int have) const {
        const GridDesc& G = P.field.grid;
        if (have >= G.nt) return INFINITY;
        const double t_grid = G.lo[0] + (double) (have - 1) * G.spacing[0];  // frame index floor(pos) <= have - 2 below it
        const double t = (t_grid - P.field.time_origin) / P.field.scale_time;
        return t - 1.001 * cfg.max_step - 1e-9 * (fabs(t) + 1.);
    }

    int set_flow_program(const dabry_leaf* leaves, int n, const double* tables, int64_t n_tab, const dabry_wrapper* w) {
        if (n < 1 || n > DABRY_MAX_LEAVES) return fail(DABRY_ERR_UNSUPPORTED, "flow program: 1..8 leaves supported");
        memcpy(P.field.leaves, leaves, sizeof(dabry_leaf) * n);
        P.field.n_leaves = n;
        be.free(d_tables);
        d_tables = nullptr;
        if (n_tab > 0) {
            d_tables = (double*) be.alloc(sizeof(double) * n_tab);
            be.upload(d_tables, tables, sizeof(double) * n_tab);
        }
        P.field.tables = d_tables;
        P.field.kind = FIELD_PROGRAM;
        set_wrapper(w);
        return check();
    }

    int set_obstacles(const dabry_obstacle* obs, int n) {
        if (n < 0 || n > DABRY_MAX_OBSTACLES) return fail(DABRY_ERR_UNSUPPORTED, "at most 8 obstacles");
        memcpy(P.obstacles, obs, sizeof(dabry_obstacle) * n);
        P.n_obstacles = n;
        return DABRY_OK;
    }

    int set_penalty(const dabry_penalty* pen) {
        if (pen->kind == PEN_GRID) return fail(DABRY_ERR_INVALID, "a gridded penalty is set with dabry_set_penalty_grid");
        memcpy(&P.penalty, pen, sizeof(dabry_penalty));
        return DABRY_OK;
    }

    // DiscretePenalty (penalty.py:96-134): data[nt][nx][ny] on the axes ts / xs / ys, through the WrapperPen parameters
    // of `pen` (kind is set here)
    int set_penalty_grid(const dabry_penalty* pen, const double* data, const double* ts, const double* xs,
                         const double* ys, int nt, int nx, int ny) {
        if (!data || !ts || !xs || !ys || nt < 2 || nx < 2 || ny < 2)
            return fail(DABRY_ERR_INVALID, "penalty grid: need data and three axes of at least two nodes");
        be.free(d_pen_data);
        be.free(d_pen_axes);
        const int64_t n = (int64_t) nt * nx * ny;
        d_pen_data = (double*) be.alloc(sizeof(double) * n);
        d_pen_axes = (double*) be.alloc(sizeof(double) * (nt + nx + ny));
        if (!d_pen_data || !d_pen_axes) return fail(DABRY_ERR_CAPACITY, "out of device memory (penalty grid)");
        be.upload(d_pen_data, data, sizeof(double) * n);
        be.upload(d_pen_axes, ts, sizeof(double) * nt);
        be.upload(d_pen_axes + nt, xs, sizeof(double) * nx);
        be.upload(d_pen_axes + nt + nx, ys, sizeof(double) * ny);
        memcpy(&P.penalty, pen, sizeof(dabry_penalty));
        P.penalty.kind = PEN_GRID;
        P.penalty_grid = PenaltyGrid{d_pen_data, d_pen_axes, d_pen_axes + nt, d_pen_axes + nt + nx, nt, nx, ny, 0};
        return check();
    }

    // ---- solve ----
    int setup(int variant_, const double* costates) {
        bind(false);
        if (P.field.kind < 0) return fail(DABRY_ERR_INVALID, "no flow field set");
        // theta_0 shard of this handle: contiguous [root_offset, root_offset + root_count) (shard_block == 0), or
        // block-cyclic: blocks of shard_block roots, every shard_world-th block of the ring starting at root_offset
        shard_b = cfg.shard_block > 0 ? cfg.shard_block : cfg.root_count;
        const int64_t world = cfg.shard_block > 0 ? cfg.shard_world : 1;
        if (cfg.root_count < 1 || cfg.root_offset < 0 || shard_b < 1 || cfg.root_count % shard_b != 0 || world < 1)
            return fail(DABRY_ERR_INVALID, "bad root range");
        n_blocks = cfg.root_count / shard_b;
        if (cfg.root_offset + (n_blocks - 1) * world * shard_b + shard_b > cfg.n_roots)
            return fail(DABRY_ERR_INVALID, "bad root range");
        variant = variant_;
        C.follow_obstacles = variant == DABRY_SIMPLE ? 0 : 1;
        has_ghost = cfg.root_count < cfg.n_roots ? 1 : 0;
        const int64_t n_ghosts = has_ghost ? n_blocks : 0;
        const int64_t n0 = cfg.root_count + n_ghosts;
        int64_t cap = cfg.max_sites > 0 ? cfg.max_sites : 2 * n0 + 1024;
        if (cap < n0) cap = n0;
        n_sites = 0;
        if (S.capacity >= cap && S.n_time == cfg.n_time) {
            // same handle solved again: keep the store, restore the NaN controls / empty neighbour links
            const int64_t cells = (int64_t) cfg.n_time * S.capacity;
            // all-ones bytes: a NaN in every control slot, -1 (None) in every neighbour link - one memset each
            be.fill_ones(S.ctrl, sizeof(Control) * cells);
            be.fill_ones(S.next_nb, sizeof(int32_t) * cells);
        } else {
            free_store(S);
            memset(&S, 0, sizeof(S));
            if (int rc = reserve(cap)) return rc;
        }
        double* d_cost = nullptr;
        if (costates) {
            d_cost = (double*) be.alloc(sizeof(double) * 2 * cfg.root_count);
            be.upload(d_cost, costates, sizeof(double) * 2 * cfg.root_count);
        }
        InitRootsPhase ph{S, d_cost, cfg.x_init[0], cfg.x_init[1], cfg.t_init, 2 * M_PI / (double) cfg.n_roots,
                          cfg.root_offset, cfg.root_count, cfg.n_roots, shard_b, world * shard_b, cfg.max_depth,
                          has_ghost};
        be.launch(cfg.root_count, ph);
        if (has_ghost) be.launch(n_blocks, GhostInitPhase{ph});  // placeholders until dabry_import_boundary
        be.sync();
        be.free(d_cost);
        n_sites = n0;
        shoot_begin = 0;
        shoot_end = (int32_t) cfg.root_count;
        round = 0;
        i_subframe = 0;
        trim_stage = 0;
        solution_site = -1;
        solution_cost = NAN;
        arrivals_ready = false;
        last_new = 0;
        be.zero(d_counters, sizeof(Counters));
        stats.rounds = 0;
        stats.ms_integrate = stats.ms_obstacle = stats.ms_resample = stats.ms_trim = stats.ms_arrival = stats.ms_total = 0.;
        stats.ms_comm = 0.;
        be.span_collect();
        return check();
    }

    int n_subframes() const {  // solver_ef.py:784-786
        const int q = cfg.n_index_per_subframe;
        return cfg.n_time / q + (cfg.n_time % q > 0 ? 1 : 0);
    }

    int step_integrate() {
        bind(false);
        if (variant == DABRY_TRIMMING) return fail(DABRY_ERR_INVALID, "step_integrate: resampling/simple only");
        if (shoot_end > shoot_begin) {
            if (int rc = integrate(nullptr, shoot_begin, shoot_end - shoot_begin, cfg.n_time - 1)) return rc;
        }
        return check();
    }

    int step_resample() {
        bind();
        shoot_begin = shoot_end = (int32_t) n_sites;
        if (variant == DABRY_RESAMPLING) {
            if (int rc = resample(cfg.n_time - 1, 0)) return rc;
            be.tic();
            if (C.ff_max_norm > 1e-8) {  // not np.isclose(max_norm, 0) (solver_ef.py:595)
                TrimDistancePhase td{C, S};
                be.launch(n_sites, td);
            }
            stats.ms_resample += be.toc();
        }
        ++round;
        stats.rounds = round;
        return check();
    }

    int step() {
        if (variant == DABRY_TRIMMING) return step_trimming();
        if (int rc = step_integrate()) return rc;
        return step_resample();
    }

    // SolverEFTrimming.step (solver_ef.py:802-817)
    // SolverEFTrimming.step (solver_ef.py:802-817) in three parts, so that a sharded run can exchange the ring
    // boundary after (1) and all_reduce(MIN) the cost map after (2):
    //   (1) integrate the sites of sites_valid[i_subframe] to the end of the sub-frame
    //   (2) fixed point of compute_new_sites / integrate the new sites; then the local cost map of trim()
    //   (3) look the open sites up in the (possibly globally reduced) cost map, close the dominated ones
    int trim_target() const {
        return imin((i_subframe + 1) * cfg.n_index_per_subframe, cfg.n_time) - 1;  // index_t_next_subframe - 1
    }

    int step_trim_integrate() {
        if (variant != DABRY_TRIMMING) return fail(DABRY_ERR_INVALID, "not a trimming solve");
        bind(false);
        if (int rc = ensure_aux(n_sites)) return rc;
        be.tic();
        ValidFlagPhase vf{S, i_subframe, d_request};
        be.launch(n_sites, vf);
        const int64_t n_list = be.scan_requests(d_request, n_sites, d_offset, d_list);
        stats.ms_trim += be.toc();
        if (n_list > 0)
            if (int rc = integrate(d_list, 0, n_list, trim_target())) return rc;
        trim_stage = 1;
        return check();
    }

    int step_trim_refine() {
        if (variant != DABRY_TRIMMING || trim_stage != 1) return fail(DABRY_ERR_INVALID, "call step_trim_integrate first");
        bind();
        const int target = trim_target();
        for (;;) {
            if (int rc = resample(target, i_subframe)) return rc;
            const int64_t n_new = shoot_end - shoot_begin;
            if (n_new == 0) break;
            if (int rc = integrate(nullptr, shoot_begin, n_new, target)) return rc;
        }
        be.tic();
        const int q = cfg.n_index_per_subframe;
        const int start_cm = i_subframe + 1 - cfg.trimming_band_width;
        // self.sites_valid[start_cm : i_subframe + 1] with Python slice semantics on a list of n_subframes + 1 sets:
        // a negative start counts from the END of the list, which leaves the slice empty for the first sub-frames
        const int len = n_subframes() + 1;
        const int a = start_cm < 0 ? imax(start_cm + len, 0) : imin(start_cm, len);
        const int b = imin(i_subframe + 1, len);
        if (int rc = build_cost_map(start_cm * q, target, 0, i_subframe, a, b)) return rc;
        stats.ms_trim += be.toc();
        trim_stage = 2;
        return check();
    }

    // on_device: `map` is a device pointer on this handle's device (an NCCL buffer of the caller, synchronised by the
    // caller before the call); the copy is ordered on the handle's stream and complete on return
    int set_cost_map(const double* map, bool on_device = false) {
        const size_t bytes = sizeof(double) * C.cm_nx * C.cm_ny;
        if (on_device) {
            be.copy2d(d_map, bytes, map, bytes, bytes, 1);
            be.sync();
        } else {
            be.upload(d_map, map, bytes);
        }
        return check();
    }

    int step_trim_close() {
        if (variant != DABRY_TRIMMING || trim_stage != 2) return fail(DABRY_ERR_INVALID, "call step_trim_refine first");
        bind();
        be.tic();
        TrimPhase tp{C, S, d_map, i_subframe, trim_target()};
        be.launch(n_sites, tp);
        stats.ms_trim += be.toc();
        trim_stage = 0;
        ++i_subframe;
        ++round;
        stats.rounds = round;
        return check();
    }

    int step_trimming() {
        if (int rc = step_trim_integrate()) return rc;
        if (int rc = step_trim_refine()) return rc;
        return step_trim_close();
    }

    int finish() {
        bind();
        if (int rc = ensure_aux(n_sites)) return rc;
        be.tic();
        ArrivalPhase ap{S, d_arr_index, d_arr_cost, 0};
        be.launch(n_sites, ap);
        if (verify_arrivals) {
            // DABRY_B200_VERIFY_ARRIVALS=1 (the test suites set it): the running arrival records against a fresh scan of
            // every trajectory, the way check_solutions reads them (solver_ef.py:667-685)
            std::vector<int32_t> idx_a(n_sites), idx_b(n_sites);
            std::vector<double> cost_a(n_sites), cost_b(n_sites);
            be.download(idx_a.data(), d_arr_index, sizeof(int32_t) * n_sites);
            be.download(cost_a.data(), d_arr_cost, sizeof(double) * n_sites);
            ArrivalPhase scan{S, d_arr_index, d_arr_cost, 1};
            be.launch(n_sites, scan);
            be.download(idx_b.data(), d_arr_index, sizeof(int32_t) * n_sites);
            be.download(cost_b.data(), d_arr_cost, sizeof(double) * n_sites);
            for (int64_t i = 0; i < n_sites; ++i)
                if (idx_a[i] != idx_b[i] || !(cost_a[i] == cost_b[i]))
                    return fail(DABRY_ERR_INVALID, "arrival records differ from a scan of the trajectories at slot " +
                                                       std::to_string(i));
        }
        // warp-level min reduction of the arrival cost, then of the slot among the minimisers (lowest slot wins)
        solution_cost = be.min_f64(d_arr_cost, n_sites);
        solution_site = -1;
        if (solution_cost < INFINITY) solution_site = be.first_equal_f64(d_arr_cost, n_sites, solution_cost);
        else solution_cost = NAN;
        arrivals_ready = true;
        stats.ms_arrival += be.toc();
        return check();
    }

    int solve(int variant_, const double* costates) {
        be.tic_total();
        if (int rc = setup(variant_, costates)) return rc;
        if (variant == DABRY_TRIMMING) {
            const int n = n_subframes();
            for (int i = 0; i < n; ++i)
                if (int rc = step_trimming()) return rc;
        } else {
            const int n = variant == DABRY_SIMPLE ? 1 : cfg.max_depth;
            for (int i = 0; i < n; ++i) {
                if (int rc = step()) return rc;
                // the remaining rounds of the reference loop (solver_ef.py:604-605) are no-ops once a round
                // neither integrates nor creates a site: compute_new_sites and trim_distance are idempotent
                if (shoot_end == shoot_begin) break;
            }
        }
        int rc = finish();
        stats.ms_total = be.toc_total();
        return rc;
    }

    // ---- results ----
    int get_sites(dabry_site_record* out, int64_t first, int64_t count) {
        if (first < 0 || count < 0 || first + count > n_sites) return fail(DABRY_ERR_INVALID, "site range");
        if (count == 0) return DABRY_OK;
        dabry_site_record* d = (dabry_site_record*) scratch(sizeof(dabry_site_record) * count);
        if (!d) return fail(DABRY_ERR_CAPACITY, "out of device memory");
        GatherRecordPhase ph{S, first, d};
        be.launch(count, ph);
        be.download(out, d, sizeof(dabry_site_record) * count);
        return check();
    }

    int get_trajs(int64_t first, int64_t count, double* times, double* states, double* costates, double* controls,
                  double* cost) {
        if (first < 0 || count < 0 || first + count > n_sites) return fail(DABRY_ERR_INVALID, "site range");
        if (count == 0) return DABRY_OK;
        const int64_t n = count * cfg.n_time;
        double* d = (double*) scratch(sizeof(double) * n * 8);
        if (!d) return fail(DABRY_ERR_CAPACITY, "out of device memory");
        GatherTrajPhase ph{S, first, cfg.n_time, times ? d : nullptr, states ? d + n : nullptr,
                           costates ? d + 3 * n : nullptr, controls ? d + 5 * n : nullptr, cost ? d + 7 * n : nullptr};
        be.launch(n, ph);
        if (times) be.download(times, d, sizeof(double) * n);
        if (states) be.download(states, d + n, sizeof(double) * 2 * n);
        if (costates) be.download(costates, d + 3 * n, sizeof(double) * 2 * n);
        if (controls) be.download(controls, d + 5 * n, sizeof(double) * 2 * n);
        if (cost) be.download(cost, d + 7 * n, sizeof(double) * n);
        return check();
    }

    int get_next_nb(int64_t first, int64_t count, int32_t* out) {
        if (first < 0 || count < 0 || first + count > n_sites) return fail(DABRY_ERR_INVALID, "site range");
        if (count == 0) return DABRY_OK;
        const int64_t n = count * cfg.n_time;
        int32_t* d = (int32_t*) scratch(sizeof(int32_t) * n);
        if (!d) return fail(DABRY_ERR_CAPACITY, "out of device memory");
        GatherNbPhase ph{S, first, cfg.n_time, d};
        be.launch(n, ph);
        be.download(out, d, sizeof(int32_t) * n);
        return check();
    }

    int get_arrivals(int64_t first, int64_t count, int32_t* index, double* cost) {
        if (!arrivals_ready) return fail(DABRY_ERR_INVALID, "call dabry_finish first");
        if (first < 0 || count < 0 || first + count > n_sites) return fail(DABRY_ERR_INVALID, "site range");
        if (index) be.download(index, d_arr_index + first, sizeof(int32_t) * count);
        if (cost) be.download(cost, d_arr_cost + first, sizeof(double) * count);
        return check();
    }

    // slots of the sites whose arrival cost is not above `bound`, ascending (order-preserving compaction on the device):
    // what check_solutions keeps, without reading the per-site arrival arrays back (12 B per site)
    int get_solution_sites(double bound, int32_t* slots, int64_t cap, int64_t* count) {
        if (!arrivals_ready) return fail(DABRY_ERR_INVALID, "call dabry_finish first");
        *count = 0;
        if (n_sites == 0) return DABRY_OK;
        be.launch(n_sites, SolutionFlagPhase{d_arr_index, d_arr_cost, bound, d_request});
        const int64_t n = be.scan_requests(d_request, n_sites, d_offset, d_list);
        *count = n;
        if (slots && n > 0 && n <= cap) be.download(slots, d_list, sizeof(int32_t) * n);
        return check();
    }

    int get_cost_map(int full, double* out, bool on_device = false) {
        bind();
        if (full) {
            if (int rc = build_cost_map(0, cfg.n_time - 1, 1, 0, 0, 0)) return rc;
        }
        const size_t bytes = sizeof(double) * C.cm_nx * C.cm_ny;
        if (on_device) {
            be.copy2d(out, bytes, d_map, bytes, bytes, 1);
            be.sync();
        } else {
            be.download(out, d_map, bytes);
        }
        return check();
    }

    int get_stats(dabry_stats* out) {
        Counters c;
        be.download(&c, d_counters, sizeof(c));
        stats.rk_steps = c.rk_steps;
        stats.ivp_calls = c.ivp_calls;
        stats.rk_steps_free = c.rk_steps_free;
        stats.rhs_evals = cfg.integrator == DABRY_RK4_FIXED ? 4 * c.rk_steps + c.ivp_calls : 6 * c.rk_steps + 2 * c.ivp_calls;
        stats.kernel_launches = be.launches();
        stats.ms_comm += be.span_collect();
        stats.n_sites = n_sites;  // device slots, including the ghost slot of a sharded ring (record flag bit 0)
        stats.capacity = S.capacity;
        *out = stats;
        return check();
    }

    int64_t num_sites() const { return n_sites; }
    double flow_max_norm = 0.;  // max |w| over the grid nodes (0 for analytic fields)
    int64_t solution() const { return solution_site; }
    double solution_cost_value() const { return solution_cost; }

    // ---- multi-GPU boundary ----
    // one packet per local block: export = the FIRST root of every block, import = the ghost of every block
    int64_t boundary_doubles() const { return (int64_t) cfg.n_time * 7 + DABRY_BOUNDARY_HEADER; }
    int64_t boundary_count() const { return n_blocks; }
    // on_device: `packed` is a device buffer of the caller on this handle's device (the tensor NCCL sends / received
    // into); the packets are written / read in place, no host staging.  The call returns with the stream drained.
    int export_boundary(double* packed, bool on_device = false) {
        const int64_t n = boundary_doubles() * n_blocks;
        double* d = on_device ? packed : (double*) scratch(sizeof(double) * n);
        if (!d) return fail(DABRY_ERR_CAPACITY, "out of device memory");
        BoundaryPhase ph{S, 0, cfg.n_time, cfg.root_count, shard_b, boundary_doubles(), d};
        be.launch(n_blocks * cfg.n_time, ph);
        if (on_device) be.sync();
        else be.download(packed, d, sizeof(double) * n);
        return check();
    }
    int import_boundary(const double* packed, bool on_device = false) {
        if (!has_ghost) return fail(DABRY_ERR_INVALID, "this handle owns the whole ring: no ghost site");
        const int64_t n = boundary_doubles() * n_blocks;
        double* d = on_device ? const_cast<double*>(packed) : (double*) scratch(sizeof(double) * n);
        if (!d) return fail(DABRY_ERR_CAPACITY, "out of device memory");
        if (!on_device) be.upload(d, packed, sizeof(double) * n);
        BoundaryPhase ph{S, 1, cfg.n_time, cfg.root_count, shard_b, boundary_doubles(), d, 0, n_blocks};
        be.launch(n_blocks * cfg.n_time, ph);
        if (on_device) be.sync();
        return check();
    }

    // ---- in-library collectives: everything below is enqueued on the solver's stream, no host round trip
    typedef typename B::Comm Comm;
    Comm* comm = nullptr;  // not owned (dabry_attach_comm)
    int comm_world() const { return comm ? B::comm_world(comm) : 1; }
    int comm_rank() const { return comm ? B::comm_rank(comm) : 0; }

    // ring boundary: the first root of every local block -> the previous rank, whose ghosts they become
    int exchange_boundary() {
        if (comm_world() <= 1) return DABRY_OK;
        if (!has_ghost) return fail(DABRY_ERR_INVALID, "sharded solve: this handle owns the whole ring");
        const int64_t n = boundary_doubles() * n_blocks;
        if (n > xchg_doubles) {
            be.free(d_send);
            be.free(d_recv);
            d_send = (double*) be.alloc(sizeof(double) * n);
            d_recv = (double*) be.alloc(sizeof(double) * n);
            if (!d_send || !d_recv) return fail(DABRY_ERR_CAPACITY, "out of device memory (boundary buffers)");
            xchg_doubles = n;
        }
        be.span_begin();  // timed by events that are read at dabry_get_stats: nothing here waits for the device
        BoundaryPhase ex{S, 0, cfg.n_time, cfg.root_count, shard_b, boundary_doubles(), d_send, 0, n_blocks};
        be.launch(n_blocks * cfg.n_time, ex);
        be.comm_ring_shift(comm, d_send, d_recv, n);
        const int64_t shift = comm_rank() == comm_world() - 1 ? 1 : 0;
        BoundaryPhase im{S, 1, cfg.n_time, cfg.root_count, shard_b, boundary_doubles(), d_recv, shift, n_blocks};
        be.launch(n_blocks * cfg.n_time, im);
        be.span_end();
        return check();
    }

    // SolverEFResampling.solve / SolverEFTrimming.solve over all ranks of the attached communicator: the ring of roots
    // is sharded by theta_0 (config.root_offset / root_count / shard_block), the ring boundary travels after the roots
    // are integrated (Trimming: every sub-frame), the trimming cost map is all-reduced (min) in place, and the arrival
    // records of all ranks are gathered at the end (global_cost / global_owner / global_rk_steps).
    int solve_sharded(int variant_, const double* costates) {
        be.tic_total();
        if (int rc = setup(variant_, costates)) return rc;
        const bool multi = comm_world() > 1;
        if (multi && !has_ghost) return fail(DABRY_ERR_INVALID, "sharded solve: config.root_count must be a shard of n_roots");
        if (variant == DABRY_TRIMMING) {
            const int n = n_subframes();
            for (int i = 0; i < n; ++i) {
                if (int rc = step_trim_integrate()) return rc;
                if (multi)
                    if (int rc = exchange_boundary()) return rc;
                if (int rc = step_trim_refine()) return rc;
                if (multi) {
                    be.span_begin();
                    be.comm_allreduce_min_f64(comm, d_map, (int64_t) C.cm_nx * C.cm_ny);
                    be.span_end();
                }
                if (int rc = step_trim_close()) return rc;
            }
        } else {
            const int n = variant == DABRY_SIMPLE ? 1 : cfg.max_depth;
            for (int i = 0; i < n; ++i) {
                if (int rc = step_integrate()) return rc;
                if (multi && i == 0)
                    if (int rc = exchange_boundary()) return rc;
                if (int rc = step_resample()) return rc;
                // every rank runs the same number of rounds only where a collective follows; after round 0 none does,
                // so a rank whose round neither integrated nor created a site can stop on its own (see solve())
                if (shoot_end == shoot_begin) break;
            }
        }
        if (int rc = finish()) return rc;
        // arrival-time argmin over ranks (check_solutions, solver_ef.py:667-685) + the totals of the throughput metric
        Counters c;
        be.download(&c, d_counters, sizeof(c));
        const int world = comm_world(), rank = comm_rank();
        std::vector<double> all(4 * (size_t) world, 0.);
        double mine[4] = {solution_cost == solution_cost ? solution_cost : INFINITY, (double) c.rk_steps, (double) n_sites, 0.};
        if (multi) {
            double* d = (double*) scratch(sizeof(double) * 4 * (world + 1));
            if (!d) return fail(DABRY_ERR_CAPACITY, "out of device memory");
            be.upload(d + 4 * world, mine, sizeof(mine));
            be.span_begin();
            be.comm_allgather(comm, d + 4 * world, d, sizeof(mine));
            be.span_end();
            be.download(all.data(), d, sizeof(double) * 4 * world);
        } else {
            memcpy(all.data(), mine, sizeof(mine));
        }
        global_cost = INFINITY;
        global_owner = -1;
        global_rk_steps = 0;
        for (int r = 0; r < world; ++r) {
            if (all[4 * r] < global_cost) {  // ties: lowest rank
                global_cost = all[4 * r];
                global_owner = r;
            }
            global_rk_steps += (uint64_t) all[4 * r + 1];
        }
        if (!(global_cost < INFINITY)) global_cost = NAN;
        (void) rank;
        stats.ms_total = be.toc_total();
        return check();
    }
    double global_cost = NAN;
    int32_t global_owner = -1;
    uint64_t global_rk_steps = 0;

    // NavigationProblem.auto_time_upper_bound (problem.py:193-239): out[0] = time_orthodromic, out[1] = time_htarget
    int time_upper_bound(const double* x_init, double t_start, double t_end, double* out) {
        bind();
        if (P.field.kind < 0) return fail(DABRY_ERR_INVALID, "no flow field set");
        double* d = (double*) scratch(sizeof(double) * 2);
        if (!d) return fail(DABRY_ERR_CAPACITY, "out of device memory");
        const double scipy_rtol = P.rtol, scipy_atol = P.atol;
        (void) scipy_rtol;
        (void) scipy_atol;
        if (P.field.kind == FIELD_GRID) be.launch(2, TimeBoundPhase<FIELD_GRID>{x_init[0], x_init[1], t_start, t_end, d, d_counters});
        else if (P.field.kind == FIELD_GRID_F32) be.launch(2, TimeBoundPhase<FIELD_GRID_F32>{x_init[0], x_init[1], t_start, t_end, d, d_counters});
        else be.launch(2, TimeBoundPhase<FIELD_PROGRAM>{x_init[0], x_init[1], t_start, t_end, d, d_counters});
        be.download(out, d, sizeof(double) * 2);
        return check();
    }

    // ---- evaluation hooks ----
    int eval_rhs(int64_t n, const double* t, const double* y, double* dy) {
        bind();
        if (P.field.kind < 0) return fail(DABRY_ERR_INVALID, "no flow field set");
        double* d = (double*) scratch(sizeof(double) * n * 9);
        if (!d) return fail(DABRY_ERR_CAPACITY, "out of device memory");
        be.upload(d, t, sizeof(double) * n);
        be.upload(d + n, y, sizeof(double) * 4 * n);
        if (P.field.kind == FIELD_GRID) be.launch(n, EvalRhsPhase<FIELD_GRID>{d, d + n, d + 5 * n});
        else if (P.field.kind == FIELD_GRID_F32) be.launch(n, EvalRhsPhase<FIELD_GRID_F32>{d, d + n, d + 5 * n});
        else be.launch(n, EvalRhsPhase<FIELD_PROGRAM>{d, d + n, d + 5 * n});
        be.download(dy, d + 5 * n, sizeof(double) * 4 * n);
        return check();
    }
    int eval_flow(int64_t n, const double* t, const double* x, double* w, double* jac) {
        bind();
        if (P.field.kind < 0) return fail(DABRY_ERR_INVALID, "no flow field set");
        double* d = (double*) scratch(sizeof(double) * n * 9);
        if (!d) return fail(DABRY_ERR_CAPACITY, "out of device memory");
        be.upload(d, t, sizeof(double) * n);
        be.upload(d + n, x, sizeof(double) * 2 * n);
        if (P.field.kind == FIELD_GRID) be.launch(n, EvalFlowPhase<FIELD_GRID>{d, d + n, d + 3 * n, d + 5 * n});
        else if (P.field.kind == FIELD_GRID_F32) be.launch(n, EvalFlowPhase<FIELD_GRID_F32>{d, d + n, d + 3 * n, d + 5 * n});
        else be.launch(n, EvalFlowPhase<FIELD_PROGRAM>{d, d + n, d + 3 * n, d + 5 * n});
        be.download(w, d + 3 * n, sizeof(double) * 2 * n);
        be.download(jac, d + 5 * n, sizeof(double) * 4 * n);
        return check();
    }

    int fail(int code, const std::string& msg) {
        err = msg;
        return code;
    }

private:
    std::vector<double> times_h;
    double* d_times = nullptr;
    double* d_tables = nullptr;
    double* d_pen_data = nullptr;  // DiscretePenalty samples and axes
    double* d_pen_axes = nullptr;
    double* d_grid = nullptr;
    int32_t* d_request = nullptr;
    int32_t* d_offset = nullptr;
    int32_t* d_list = nullptr;
    int32_t* d_captured = nullptr;
    int32_t* d_cap_offset = nullptr;
    int32_t* d_cap_list = nullptr;
    int32_t* d_run_a = nullptr;
    int32_t* d_run_b = nullptr;
    int64_t work_capacity = 0;
    double* d_map = nullptr;
    int32_t* d_arr_index = nullptr;
    double* d_arr_cost = nullptr;
    Counters* d_counters = nullptr;
    double* d_send = nullptr;   // ring-boundary packets of the in-library exchange
    double* d_recv = nullptr;
    int64_t xchg_doubles = 0;
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    int64_t aux_capacity = 0;
    int64_t n_sites = 0;
    int32_t shoot_begin = 0, shoot_end = 0;
    int variant = DABRY_RESAMPLING;
    int has_ghost = 0;
    int64_t shard_b = 1, n_blocks = 1;
    int round = 0, i_subframe = 0, trim_stage = 0;
    // work items up to which the latency build of the arc kernels is launched.  The two builds differ in how ptxas
    // contracts multiply-adds (<= 1 ulp per operation); DABRY_B200_WIDE_BATCH=0 pins every batch to the lean build
    int64_t wide_batch = DABRY_WIDE_BATCH;
    int64_t spread_batch = DABRY_SPREAD_BATCH;  // <= : one extremal per warp (Spread); DABRY_B200_SPREAD_BATCH overrides
    bool verify_arrivals = false;
    // streamed grid upload (set_flow_grid / pack_groups / integrate); DABRY_B200_STREAM_UPLOAD=0 disables it
    int stream_upload = 1;  // 0: never, 1: grids of 64 MB and more, 2: every eligible grid (tests)
    bool grid_pending = false;
    int n_groups = 4, groups_packed = 0, frames_packed = 0, group_frames[kMaxGridGroups] = {};
    PackGridPhase pending_pack{};
    double* pending_values = nullptr;
    const double* pending_host = nullptr;
    int groups_issued = 0;
    unsigned long long* pending_key = nullptr;
    int free_chunk = DABRY_FREE_CHUNK;  // RK steps per launch of the free-arc kernel (0: whole arc in one launch)
    int64_t last_new = 0;
    int64_t solution_site = -1;
    double solution_cost = NAN;
    bool arrivals_ready = false;

    void set_wrapper(const dabry_wrapper* w) {
        FieldDesc& F = P.field;
        F.time_origin = w ? w->time_origin : 0.;
        F.scale_time = w ? w->scale_time : 1.;
        F.scale_length = w ? w->scale_length : 1.;
        F.wbl[0] = w ? w->bl[0] : 0.;
        F.wbl[1] = w ? w->bl[1] : 0.;
        F.scaler_speed = w ? w->scaler_speed : 1.;
        F.scaler_dspeed = w ? w->scaler_dspeed : 1.;
    }

    // need_grid = false: the caller copes with a grid that is still arriving (setup; integrate packs it in windows)
    void bind(bool need_grid = true) {
        if (need_grid) finish_grid();
        be.set_problem(P);
    }

    template <int V>
    struct Tag {
        static constexpr int value = V;
    };
    // compile-time (field family, integrator) for the two integrate kernels
    template <class Fn>
    void dispatch_field(bool rk4, Fn&& fn) {
        const int k = P.field.kind;
        if (k == FIELD_GRID) {
            if (rk4) fn(Tag<FIELD_GRID>{}, Tag<METHOD_RK4>{});
            else fn(Tag<FIELD_GRID>{}, Tag<METHOD_DOPRI5>{});
        } else if (k == FIELD_GRID_F32) {
            if (rk4) fn(Tag<FIELD_GRID_F32>{}, Tag<METHOD_RK4>{});
            else fn(Tag<FIELD_GRID_F32>{}, Tag<METHOD_DOPRI5>{});
        } else {
            if (rk4) fn(Tag<FIELD_PROGRAM>{}, Tag<METHOD_RK4>{});
            else fn(Tag<FIELD_PROGRAM>{}, Tag<METHOD_DOPRI5>{});
        }
    }

    int check() {
        if (!be.ok()) return fail(DABRY_ERR_CUDA, be.error());
        return DABRY_OK;
    }

    void* scratch(size_t bytes) {
        if (bytes > scratch_bytes) {
            be.free(d_scratch);
            d_scratch = be.alloc(bytes);
            scratch_bytes = d_scratch ? bytes : 0;
        }
        return d_scratch;
    }

    int ensure_work(int64_t n) {
        if (n <= work_capacity) return DABRY_OK;
        const int64_t cap = n + n / 2 + 1024;
        be.free(d_captured);
        be.free(d_cap_offset);
        be.free(d_cap_list);
        be.free(d_run_a);
        be.free(d_run_b);
        d_run_a = (int32_t*) be.alloc(sizeof(int32_t) * cap);
        d_run_b = (int32_t*) be.alloc(sizeof(int32_t) * cap);
        d_captured = (int32_t*) be.alloc(sizeof(int32_t) * cap);
        d_cap_offset = (int32_t*) be.alloc(sizeof(int32_t) * cap);
        d_cap_list = (int32_t*) be.alloc(sizeof(int32_t) * cap);
        if (!d_captured || !d_cap_offset || !d_cap_list || !d_run_a || !d_run_b) return fail(DABRY_ERR_CAPACITY, "out of device memory (work)");
        work_capacity = cap;
        return DABRY_OK;
    }

    int ensure_aux(int64_t n) {
        if (n <= aux_capacity) return DABRY_OK;
        const int64_t cap = n + n / 2 + 1024;
        be.free(d_request);
        be.free(d_offset);
        be.free(d_list);
        be.free(d_arr_index);
        be.free(d_arr_cost);
        d_request = (int32_t*) be.alloc(sizeof(int32_t) * cap);
        d_offset = (int32_t*) be.alloc(sizeof(int32_t) * cap);
        d_list = (int32_t*) be.alloc(sizeof(int32_t) * cap);
        d_arr_index = (int32_t*) be.alloc(sizeof(int32_t) * cap);
        d_arr_cost = (double*) be.alloc(sizeof(double) * cap);
        if (!d_request || !d_offset || !d_list || !d_arr_index || !d_arr_cost)
            return fail(DABRY_ERR_CAPACITY, "out of device memory (aux)");
        aux_capacity = cap;
        arrivals_ready = false;
        return DABRY_OK;
    }

    // One array of the site store during a re-pitch: where the pointer lives, element size, [rows][capacity * per] shape
    struct StoreArray {
        void** slot;
        size_t elem;
        int64_t rows, per;
        int fill_ones;  // fresh memory = all-ones bytes (NaN controls, -1 links)
        void* fresh;
    };
    template <class T>
    static StoreArray store_array(T*& p, int64_t rows, int64_t per, int fill_ones = 0) {
        return StoreArray{reinterpret_cast<void**>(&p), sizeof(T), rows, per, fill_ones, nullptr};
    }

    void free_store(SiteStore& s) {
        for (StoreArray& a : store_arrays(s)) {
            be.free((char*) *a.slot);
            *a.slot = nullptr;
        }
    }

    std::vector<StoreArray> store_arrays(SiteStore& s) {
        const int64_t T = cfg.n_time;
        return {
            store_array(s.xp, T, 1), store_array(s.ctrl, T, 1, 1), store_array(s.tt, T, 1), store_array(s.next_nb, T, 1, 1),
            store_array(s.cost0, 1, 1), store_array(s.t_enter_obs, 1, 1), store_array(s.index_t_init, 1, 1),
            store_array(s.n_samples, 1, 1), store_array(s.closure, 1, 1), store_array(s.neuter, 1, 1),
            store_array(s.index_neutered, 1, 1), store_array(s.obstacle, 1, 1), store_array(s.index_t_obs, 1, 1),
            store_array(s.trigo, 1, 1), store_array(s.check_next, 1, 1), store_array(s.depth, 1, 1),
            store_array(s.flags, 1, 1), store_array(s.parent_prev, 1, 1), store_array(s.parent_next, 1, 1),
            store_array(s.created_subframe, 1, 1), store_array(s.valid_first, 1, 1), store_array(s.valid_last, 1, 1),
            store_array(s.dyadic, 1, 1), store_array(s.n_target_events, 1, 1),
            store_array(s.t_target_events, 1, DABRY_MAX_TARGET_EVENTS), store_array(s.ivp_status, 1, 1),
            store_array(s.arc_from, 1, 1), store_array(s.enter, 1, 1), store_array(s.arr_index, 1, 1),
            store_array(s.arr_cost, 1, 1), store_array(s.carry_t, 1, 1), store_array(s.carry_h, 1, 1),
            store_array(s.carry_y, 1, 1), store_array(s.carry_f, 1, 1), store_array(s.carry_g, 1, DABRY_MAX_EVENTS),
            store_array(s.carry_eval_i, 1, 1),
        };
    }

    // Grow the site store to at least `want` slots.  Every new array is allocated FIRST; only when all allocations have
    // succeeded are the rows re-pitched and the pointers and the capacity committed, so an out-of-memory failure leaves
    // the store as it was (old pitch, old capacity).
    int reserve(int64_t want) {
        if (want <= S.capacity) return DABRY_OK;
        if (want > INT_MAX / 2) return fail(DABRY_ERR_CAPACITY, "more than 2^30 sites");
        int64_t cap = S.capacity > 0 ? S.capacity + S.capacity / 2 : want;
        if (cap < want) cap = want;
        cap = (cap + 31) / 32 * 32;  // keep rows 256-byte aligned
        const int64_t oc = S.capacity;
        std::vector<StoreArray> arrays = store_arrays(S);
        bool ok = true;
        for (StoreArray& a : arrays) {
            a.fresh = be.alloc(a.elem * a.rows * cap * a.per);
            if (!a.fresh) {
                ok = false;
                break;
            }
        }
        if (!ok) {
            for (StoreArray& a : arrays) be.free((char*) a.fresh);
            return fail(DABRY_ERR_CAPACITY, "out of device memory (site store)");
        }
        for (StoreArray& a : arrays) {
            // controls and neighbour links start as NaN / -1 everywhere (free arcs never write controls)
            if (a.fill_ones) be.fill_ones(a.fresh, a.elem * a.rows * cap * a.per);
            if (*a.slot && n_sites > 0)  // [rows][cap * per] arrays: re-pitch row by row
                be.copy2d(a.fresh, a.elem * cap * a.per, *a.slot, a.elem * oc * a.per, a.elem * n_sites * a.per, a.rows);
            be.free((char*) *a.slot);
            *a.slot = a.fresh;
        }
        S.capacity = (int32_t) cap;
        S.n_time = cfg.n_time;
        S.t_init = cfg.t_init;
        S.times = d_times;
        be.sync();
        return check();
    }

    int integrate(const int32_t* list, int32_t base, int64_t n, int index_t) {
        if (int rc = ensure_work(n)) return rc;
        be.tic();
        const bool rk4 = cfg.integrator == DABRY_RK4_FIXED;
        // chunks of `free_chunk` RK steps; between chunks the extremals still running are re-packed into a dense
        // list (ballot/prefix-sum compaction) so that warps stay full after their neighbours were captured
        const int32_t* cur = list;
        int32_t cur_base = base;
        int64_t n_cur = n;
        int flip = 0;
        if (grid_pending) {
            // the grid is still arriving: one window of free arcs per group of frames, every arc parked (IvpCarry) at the
            // time the frames that are on the device run out, while the copy stream brings the next group
            bool first = true;
            while (grid_pending) {
                const int g = groups_packed;
                pack_groups(g);
                const double t_park = park_time(group_frames[g]);
                dispatch_field(rk4, [&](auto fk, auto method) {
                    FreeArcPhase<decltype(fk)::value, decltype(method)::value, true> ph{
                        S, list, base, index_t, 0x7fffffff, d_counters, d_captured};
                    ph.t_park = t_park;
                    ph.resume_only = first ? 0 : 1;
                    be.launch(n, ph);
                });
                first = false;
            }
            n_cur = 0;  // the last window ran every arc to its end
        }
        while (n_cur > 0) {
            dispatch_field(rk4, [&](auto fk, auto method) {
                constexpr int kField = decltype(fk)::value, kMethod = decltype(method)::value;
                if (free_chunk > 0)
                    be.launch(n_cur, FreeArcPhase<kField, kMethod, true>{S, cur, cur_base, index_t, free_chunk,
                                                                         d_counters, d_captured});
                else if (n_cur <= spread_batch)
                    be.launch(n_cur * 32, Spread<FreeArcPhase<kField, kMethod, false, false>>{
                                              {S, cur, cur_base, index_t, 0, d_counters, d_captured}});
                else if (n_cur <= wide_batch)
                    be.launch(n_cur, FreeArcPhase<kField, kMethod, false, false>{S, cur, cur_base, index_t, 0,
                                                                                 d_counters, d_captured});
                else
                    be.launch(n_cur, FreeArcPhase<kField, kMethod, false>{S, cur, cur_base, index_t, 0, d_counters,
                                                                          d_captured});
            });
            if (free_chunk <= 0) break;
            int32_t* next = flip ? d_run_b : d_run_a;
            n_cur = be.scan_requests(d_captured, n_cur, d_cap_offset, next, cur, cur_base);
            cur = next;
            cur_base = 0;
            flip ^= 1;
        }
        be.launch(n, CapturedFlagPhase{S, list, base, d_captured});
        stats.ms_integrate += be.toc();
        if (C.follow_obstacles) {
            be.tic();
            const int64_t n_cap = be.scan_requests(d_captured, n, d_cap_offset, d_cap_list);
            if (n_cap > 0) {
                dispatch_field(rk4, [&](auto fk, auto method) {
                    constexpr int kField = decltype(fk)::value, kMethod = decltype(method)::value;
                    if (n_cap <= spread_batch)
                        be.launch(n_cap * 32, Spread<ObstacleArcPhase<kField, kMethod, false>>{
                                                  {S, list, d_cap_list, base, index_t, d_counters}});
                    else if (n_cap <= wide_batch)
                        be.launch(n_cap, ObstacleArcPhase<kField, kMethod, false>{S, list, d_cap_list, base, index_t,
                                                                                  d_counters});
                    else
                        be.launch(n_cap, ObstacleArcPhase<kField, kMethod>{S, list, d_cap_list, base, index_t,
                                                                           d_counters});
                });
            }
            stats.ms_obstacle += be.toc();
        }
        ConnectPhase cp{S, list, base};
        be.launch(n, cp);
        return check();
    }

    // compute_new_sites(index_t_hi) (solver_ef.py:630-665); new sites become [shoot_begin, shoot_end)
    int resample(int index_t_hi, int subframe) {
        if (int rc = ensure_aux(n_sites)) return rc;
        be.tic();
        ResamplePhase rp{C, S, index_t_hi, d_request};
        be.launch(n_sites, rp);
        const int64_t n_new = be.scan_requests(d_request, n_sites, d_offset, nullptr);
        if (n_new > 0) {
            if (int rc = reserve(n_sites + n_new)) return rc;
            SpawnPhase sp{C, S, d_request, d_offset, (int32_t) n_sites, subframe};
            be.launch(n_sites, sp);
        }
        shoot_begin = (int32_t) n_sites;
        n_sites += n_new;
        shoot_end = (int32_t) n_sites;
        last_new = n_new;
        arrivals_ready = false;
        stats.ms_resample += be.toc();
        return check();
    }

    int build_cost_map(int k_lo, int k_hi, int all_sites, int subframe, int va, int vb) {
        be.launch((int64_t) C.cm_nx * C.cm_ny, FillF64Phase{d_map, INFINITY});
        const int lo = imax(k_lo, 0);
        if (k_hi >= lo && n_sites > 0) {
            RasterPhase rp{C, S, d_map, (int32_t) n_sites, lo, k_hi, subframe, va, vb, all_sites};
            be.launch(n_sites * (int64_t) (k_hi - lo + 1), rp);
        }
        return check();
    }
};

}  // namespace dabry
