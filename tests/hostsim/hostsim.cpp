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

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>

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

// DABRY_HOSTSIM_TRACE_RK=<file>: one "n t h error_norm" line per RK step attempt (accept / reject analysis)
static FILE* hostsim_rk_trace_file() {
    static FILE* f = getenv("DABRY_HOSTSIM_TRACE_RK") ? fopen(getenv("DABRY_HOSTSIM_TRACE_RK"), "w") : nullptr;
    return f;
}
#define DABRY_TRACE_RK_ATTEMPT(n, t, h, error_norm)                                                    \
    do {                                                                                               \
        if (FILE* tf_ = hostsim_rk_trace_file()) fprintf(tf_, "%d %.17g %.17g %.17g\n", (int) (n), (double) (t), (double) (h), (double) (error_norm)); \
    } while (0)

// DABRY_HOSTSIM_TRACE_CELL=1: how often consecutive grid samples of the host thread fall into the same cell and frame
// pair (what a per-thread cell cache would save); printed at exit
struct HostsimCellStats {
    long long samples = 0, same = 0, same_xy = 0;
    int frame = -9, i0 = -9, j0 = -9;
    bool on = getenv("DABRY_HOSTSIM_TRACE_CELL") != nullptr;
    ~HostsimCellStats() {
        if (on && samples) fprintf(stderr, "hostsim cells: %lld samples, same cell+frames as the previous sample %.4f, same cell %.4f\n",
                                   samples, (double) same / samples, (double) same_xy / samples);
    }
};
static HostsimCellStats g_cell_stats;
#define DABRY_TRACE_CELL(f_, i_, j_)                                                                            \
    do {                                                                                                        \
        if (g_cell_stats.on) {                                                                                  \
            ++g_cell_stats.samples;                                                                             \
            if (g_cell_stats.i0 == (i_) && g_cell_stats.j0 == (j_)) {                                           \
                ++g_cell_stats.same_xy;                                                                         \
                if (g_cell_stats.frame == (f_)) ++g_cell_stats.same;                                            \
            }                                                                                                   \
            g_cell_stats.frame = (f_); g_cell_stats.i0 = (i_); g_cell_stats.j0 = (j_);                          \
        }                                                                                                       \
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
    void upload_async(void* d, const void* h, size_t n) { memcpy(d, h, n); }
    struct CallGuard {  // one thread, one "device": nothing to serialise
        explicit CallGuard(SerialBackend&) {}
    };
    template <class Fn>
    static void destroy_locked(SerialBackend&, Fn&& destroy) { destroy(); }
    void span_begin() { span_t0_ = now(); }
    void span_end() { span_ms_ += now() - span_t0_; }
    double span_collect() {
        const double v = span_ms_;
        span_ms_ = 0.;
        return v;
    }

    // ---- collectives of the CPU tier: the ranks are processes of one host, the transport is a POSIX shared-memory
    // segment named after the unique id, one mailbox per ordered pair of ranks (test double of the NCCL communicator
    // of the product library; same call sequence, same engine code above it)
    static constexpr size_t kMailBytes = (size_t) 1 << 20;
    static constexpr int kMaxWorld = 8;
    struct Mailbox {
        std::atomic<uint64_t> written, read;
        uint64_t bytes;
        char data[kMailBytes];
    };
    struct Segment {
        std::atomic<int> ready, attached;
        Mailbox box[kMaxWorld][kMaxWorld];  // [src][dst]
    };
    struct Comm {
        Segment* seg = nullptr;
        int rank = 0, world = 1;
        std::string name;
    };
    static int comm_rank(const Comm* c) { return c->rank; }
    static int comm_world(const Comm* c) { return c->world; }
    static bool comm_unique_id(void* id, std::string&) {
        unsigned char* b = (unsigned char*) id;
        memset(b, 0, 128);
        FILE* f = fopen("/dev/urandom", "rb");
        if (f) {
            if (fread(b, 1, 16, f) != 16) b[0] = (unsigned char) getpid();
            fclose(f);
        }
        return true;
    }
    static Comm* comm_create(int, int rank, int world, const void* id, std::string& err) {
        if (world > kMaxWorld) {
            err = "hostsim communicator: at most 8 ranks";
            return nullptr;
        }
        char name[64] = "/dabry_hostsim_";
        const unsigned char* b = (const unsigned char*) id;
        for (int i = 0; i < 16; ++i) snprintf(name + 15 + 2 * i, 3, "%02x", b[i]);
        Comm* c = new Comm;
        c->rank = rank;
        c->world = world;
        c->name = name;
        int fd = -1;
        if (rank == 0) {
            fd = shm_open(name, O_CREAT | O_RDWR, 0600);
            if (fd >= 0 && ftruncate(fd, sizeof(Segment)) != 0) {
                close(fd);
                fd = -1;
            }
        } else {
            for (int tries = 0; tries < 6000 && fd < 0; ++tries) {
                fd = shm_open(name, O_RDWR, 0600);
                struct stat st;
                if (fd >= 0 && (fstat(fd, &st) != 0 || (size_t) st.st_size < sizeof(Segment))) {  // not sized yet
                    close(fd);
                    fd = -1;
                }
                if (fd < 0) std::this_thread::sleep_for(std::chrono::milliseconds(10));
            }
        }
        void* p = fd >= 0 ? mmap(nullptr, sizeof(Segment), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0) : MAP_FAILED;
        if (fd >= 0) close(fd);
        if (p == MAP_FAILED) {
            err = "hostsim communicator: shared memory segment unavailable";
            delete c;
            return nullptr;
        }
        c->seg = (Segment*) p;  // a fresh segment is zero-filled: counters start at 0
        if (rank == 0) c->seg->ready.store(1);
        while (c->seg->ready.load() == 0) std::this_thread::sleep_for(std::chrono::milliseconds(1));
        c->seg->attached.fetch_add(1);
        while (c->seg->attached.load() < world) std::this_thread::sleep_for(std::chrono::milliseconds(1));
        if (rank == 0) shm_unlink(name);  // everybody holds a mapping: the name can go
        return c;
    }
    static void comm_destroy(Comm* c) {
        if (!c) return;
        if (c->seg) munmap(c->seg, sizeof(Segment));
        delete c;
    }
    static void post(Mailbox& m, const char* src, size_t n) {
        while (m.written.load() != m.read.load()) std::this_thread::yield();
        memcpy(m.data, src, n);
        m.bytes = n;
        m.written.fetch_add(1);
    }
    static void take(Mailbox& m, char* dst, size_t n) {
        while (m.written.load() == m.read.load()) std::this_thread::yield();
        memcpy(dst, m.data, n);
        m.read.fetch_add(1);
    }
    // chunked so that two ranks sending to each other never wait for an empty mailbox at the same time
    static void sendrecv(Comm* c, const void* send, int to, void* recv, int from, size_t bytes) {
        for (size_t off = 0; off < bytes || off == 0; off += kMailBytes) {
            const size_t n = bytes - off < kMailBytes ? bytes - off : kMailBytes;
            post(c->seg->box[c->rank][to], (const char*) send + off, n);
            take(c->seg->box[from][c->rank], (char*) recv + off, n);
            if (bytes == 0) break;
        }
    }
    void comm_ring_shift(Comm* c, const double* send, double* recv, int64_t n) {
        sendrecv(c, send, (c->rank + c->world - 1) % c->world, recv, (c->rank + 1) % c->world, sizeof(double) * n);
    }
    void comm_allgather(Comm* c, const void* send, void* recv, size_t bytes) {
        std::string mine((const char*) send, bytes);  // `send` may alias its own slot of `recv`
        memcpy((char*) recv + c->rank * bytes, mine.data(), bytes);
        for (int d = 1; d < c->world; ++d) {
            const int to = (c->rank + d) % c->world, from = (c->rank + c->world - d) % c->world;
            sendrecv(c, mine.data(), to, (char*) recv + from * bytes, from, bytes);
        }
    }
    void comm_allreduce_min_f64(Comm* c, double* buf, int64_t n) {
        std::vector<double> all((size_t) n * c->world);
        comm_allgather(c, buf, all.data(), sizeof(double) * n);
        for (int r = 0; r < c->world; ++r)
            for (int64_t i = 0; i < n; ++i)
                if (all[(size_t) r * n + i] < buf[i]) buf[i] = all[(size_t) r * n + i];
    }
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
    double t0_ = 0., t1_ = 0., span_t0_ = 0., span_ms_ = 0.;
};

}  // namespace dabry

#define DABRY_BACKEND dabry::SerialBackend
#include "../../dabry_b200/csrc/capi.inl"
