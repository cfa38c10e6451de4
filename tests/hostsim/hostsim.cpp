// TEST INFRASTRUCTURE - not part of the dabry_b200 package and never loaded by it.
//
// libdabry_hostsim.so = the device headers of dabry_b200/csrc compiled by g++ with a serial backend behind the
// same C ABI, so that the CPU-only test tier (`pytest -m "not gpu"`, no GPU in the build container) can step
// through the very source the CUDA kernels are built from and check it against the reference's golden vectors
// before any GPU time is spent.  It shares the product's source on purpose, so it is NOT the parity oracle (that
// is oracle/ + tests/golden) and it is not a fallback: dabry_b200/_native.py only ever loads libdabry_b200.so.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>

// DABRY_HOSTSIM_TRACE=<file>: one "site attempts" line per free arc (lane-imbalance analysis, profiles/lane_imbalance.py)
#include <stdio.h>
static FILE* hostsim_trace_file() {
    static FILE* f = getenv("DABRY_HOSTSIM_TRACE") ? fopen(getenv("DABRY_HOSTSIM_TRACE"), "w") : nullptr;
    return f;
}
#define DABRY_TRACE_FREE_ARC(site, attempts)                                                    \
    do {                                                                                        \
        if (FILE* tf_ = hostsim_trace_file()) { fprintf(tf_, "%d %u\n", (int) (site), (unsigned) (attempts)); fflush(tf_); } \
    } while (0)

#include "../../dabry_b200/csrc/engine.cuh"

namespace dabry {

class SerialBackend {
public:
    static const char* name() { return "hostsim-serial"; }
    static bool host_register(void*, size_t) { return true; }
    static bool host_unregister(void*) { return true; }
    static bool release_cached(int) { return true; }
    bool open(int) { return true; }
    void activate() const {}
    bool ok() const { return true; }
    const std::string& error() const { return err_; }
    uint64_t launches() const { return launches_; }
    // fresh "device" memory is poisoned (all-ones bytes: NaN doubles, -1 ints) so that a read of something the engine
    // never wrote - e.g. a record frame of a streamed grid that has not been packed yet - changes results in the CPU tier
    void* alloc(size_t bytes) {
        void* p = malloc(bytes ? bytes : 8);
        if (p) memset(p, 0xFF, bytes ? bytes : 8);
        return p;
    }
    template <class T>
    void free(T* p) { ::free((void*) p); }
    void upload(void* d, const void* h, size_t n) { memcpy(d, h, n); }
    void download(void* h, const void* d, size_t n) { memcpy(h, d, n); }
    // streamed grid upload: the serial double copies at once; DABRY_HOSTSIM_STREAMED=1 makes every host array count as
    // page-locked so that the CPU tier runs the windowed free arcs
    void upload_chunk_async(void* d, const void* h, size_t n, int) { memcpy(d, h, n); }
    void wait_chunk(int) {}
    void sync_copies() {}
    static bool host_is_pinned(const void*) { return getenv("DABRY_HOSTSIM_STREAMED") != nullptr; }
    void zero(void* d, size_t n) { memset(d, 0, n); }
    void fill_ones(void* d, size_t n) { memset(d, 0xFF, n); }
    void copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows) {
        for (size_t r = 0; r < rows; ++r) memcpy((char*) dst + r * dpitch, (const char*) src + r * spitch, width);
    }
    void sync() {}
    void set_problem(const ProblemDev& P) { g_problem = P; }
    template <class F>
    void launch(int64_t n, const F& f) {
        ++launches_;
        for (int64_t i = 0; i < n; ++i) f(i);
    }
    int64_t scan_requests(const int32_t* request, int64_t n, int32_t* offset, int32_t* list,
                          const int32_t* values = nullptr, int32_t value_base = 0) {
        int32_t acc = 0;
        for (int64_t i = 0; i < n; ++i) {
            offset[i] = acc;
            if (request[i] >= 0) {
                if (list) list[acc] = values ? values[i] : value_base + (int32_t) i;
                ++acc;
            }
        }
        return acc;
    }
    double min_f64(const double* v, int64_t n) {
        double m = INFINITY;
        for (int64_t i = 0; i < n; ++i)
            if (v[i] < m) m = v[i];
        return m;
    }
    int64_t first_equal_f64(const double* v, int64_t n, double value) {
        for (int64_t i = 0; i < n; ++i)
            if (v[i] == value) return i;
        return -1;
    }
    void tic() { t0_ = now(); }
    double toc() { return now() - t0_; }
    void tic_total() { t1_ = now(); }
    double toc_total() { return now() - t1_; }

private:
    static double now() {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    std::string err_;
    uint64_t launches_ = 0;
    double t0_ = 0., t1_ = 0.;
};

}  // namespace dabry

#define DABRY_BACKEND dabry::SerialBackend
#include "../../dabry_b200/csrc/capi.inl"
