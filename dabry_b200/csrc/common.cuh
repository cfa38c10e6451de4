// dabry_b200 device core - common definitions.
//
// Every per-extremal routine in csrc/*.cuh is a __host__ __device__ function: the product path only ever
// runs them inside the CUDA kernels of dabry_b200.cu; tests/hostsim compiles the same headers with g++ to
// step through the identical source on the CPU for debugging (test infrastructure, not a fallback: the
// shared library has no host execution path and refuses to create a solver without a CUDA device).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DABRY_HD __host__ __device__ __forceinline__
#define DABRY_HD_NOINLINE __host__ __device__ __noinline__
#else
#define DABRY_HD inline
#define DABRY_HD_NOINLINE inline
#endif

#define DABRY_MAX_OBSTACLES 8
#define DABRY_MAX_LEAVES 8
#define DABRY_MAX_TARGET_EVENTS 4

namespace dabry {

// --- arithmetic that must not be contracted into FMAs -------------------------------------------------
// The reference differentiates penalties by central differences with dx = 1e-8 (penalty.py:39-44): one ulp
// of difference in value(x +- dx) is amplified 5e7 times, so those few operations replay numpy's
// individually rounded *, +, - (host build: -ffp-contract=off).
DABRY_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
DABRY_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
DABRY_HD double sub_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, -b);
#else
    return a - b;
#endif
}

// 1 / sqrt(x): the device's rsqrt (<= 1 ulp) instead of a square root followed by a division
DABRY_HD double inv_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1. / sqrt(x);
#endif
}

DABRY_HD double dmin(double a, double b) { return a < b ? a : b; }
DABRY_HD double dmax(double a, double b) { return a > b ? a : b; }
DABRY_HD int imin(int a, int b) { return a < b ? a : b; }
DABRY_HD int imax(int a, int b) { return a > b ? a : b; }
DABRY_HD int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

DABRY_HD double quiet_nan() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return NAN;
#endif
}

// spacing to the next representable double above |t| (numpy.nextafter(t, inf) - t for t >= 0, rk.py:119)
DABRY_HD double ulp_above(double t) {
    double a = fabs(t);
    return nextafter(a, (double) HUGE_VAL) - a;
}

enum Coords : int32_t { COORDS_CARTESIAN = 0, COORDS_GCS = 1 };

// termination flags, numerically identical to the reference enums (solver_ef.py:42-53); -1 = None
enum Closure : int32_t { CLOSURE_NONE = -1, CLOSURE_TERMINAL_INTEGRATION = 0, CLOSURE_WITHIN_OBS = 1,
                         CLOSURE_SUBOPTIMAL = 2, CLOSURE_IMPOSSIBLE_OBS_TRACKING = 3 };
enum Neuter : int32_t { NEUTER_NONE = -1, NEUTER_SELF_AND_NB_IN_OBS = 0 };

}  // namespace dabry
