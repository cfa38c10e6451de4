// libdabry_b200.so - the product library: Engine<CudaBackend> behind the C ABI of include/dabry_b200.h.
// Built for sm_100a only (nvcc -gencode arch=compute_100a,code=sm_100a).  There is no host execution path in this
// translation unit: every phase functor runs as a CUDA kernel, one thread per extremal / site / grid node.
#include <cuda_runtime.h>
#include <stdio.h>

#include <chrono>
#include <map>
#include <mutex>
#include <unordered_map>

#include "engine.cuh"
#include "nccl_dyn.cuh"

namespace dabry {

// one thread per work item; the functor (problem constants, site-store pointers) travels as a __grid_constant__
// kernel parameter so that uniform reads of the problem description hit the constant bank
template <class F, class = void>
struct MinBlocks {
    static constexpr int value = 1;
};
template <class F>
struct MinBlocks<F, decltype((void) F::kMinBlocks)> {
    static constexpr int value = F::kMinBlocks;
};

template <class F>
__global__ void __launch_bounds__(F::kThreads, MinBlocks<F>::value) phase_kernel(int64_t n, const __grid_constant__ F f) {
    const int64_t i = (int64_t) blockIdx.x * F::kThreads + threadIdx.x;
    if (i < n) f(i);
}

// ---- order-preserving stream compaction (resampling insertions, trimming survivors) -----------------------
// flags are (request >= 0).  Pass 1: warp ballot + popc gives the in-warp rank, a shuffle scan over the 32 warp
// totals gives the in-block rank.  Pass 2: one block scans the block totals.  Pass 3: add the block base and,
// optionally, scatter the surviving indices into a dense list.
constexpr int kScanBlock = 1024;

__global__ void __launch_bounds__(kScanBlock) scan_pass1(const int32_t* __restrict__ request, int64_t n,
                                                         int32_t* __restrict__ offset, int32_t* __restrict__ block_sum) {
    __shared__ int32_t warp_sum[32];
    const int64_t i = (int64_t) blockIdx.x * kScanBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool flag = i < n && request[i] >= 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, flag);
    const int rank = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_sum[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        int v = warp_sum[lane];
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        warp_sum[lane] = incl - v;  // exclusive
        if (lane == 31) block_sum[blockIdx.x] = incl;
    }
    __syncthreads();
    if (i < n) offset[i] = warp_sum[warp] + rank;
}

__global__ void __launch_bounds__(kScanBlock) scan_pass2(int32_t* __restrict__ block_sum, int64_t n_blocks,
                                                         int64_t* __restrict__ total) {
    __shared__ int32_t warp_sum[32];
    __shared__ int32_t carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += kScanBlock) {
        const int64_t i = base + threadIdx.x;
        const int v = i < n_blocks ? block_sum[i] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = warp_sum[lane];
            int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += up;
            }
            warp_sum[lane] = wi - w;
        }
        __syncthreads();
        const int excl = carry + warp_sum[warp] + incl - v;
        if (i < n_blocks) block_sum[i] = excl;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kScanBlock) scan_pass3(const int32_t* __restrict__ request, int64_t n,
                                                         int32_t* __restrict__ offset,
                                                         const int32_t* __restrict__ block_sum,
                                                         int32_t* __restrict__ list,
                                                         const int32_t* __restrict__ values, int32_t value_base) {
    const int64_t i = (int64_t) blockIdx.x * kScanBlock + threadIdx.x;
    if (i >= n) return;
    const int32_t o = offset[i] + block_sum[blockIdx.x];
    offset[i] = o;
    // the dense list holds the flagged indices, or values[i] / value_base + i when a value map is given
    if (list && request[i] >= 0) list[o] = values ? values[i] : value_base + (int32_t) i;
}

// ---- arrival-time reductions (check_solutions) --------------------------------------------------------------
__device__ __forceinline__ unsigned long long f64_order_key(double v) {  // total order of doubles as unsigned keys
    const unsigned long long b = (unsigned long long) __double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256) min_f64_kernel(const double* __restrict__ v, int64_t n,
                                                      unsigned long long* __restrict__ out_key) {
    unsigned long long best = ~0ull;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        const unsigned long long k = f64_order_key(v[i]);
        best = k < best ? k : best;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
        best = o < best ? o : best;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(out_key, best);
}

__global__ void __launch_bounds__(256) first_equal_kernel(const double* __restrict__ v, int64_t n, double value,
                                                          unsigned long long* __restrict__ out_index) {
    unsigned long long best = ~0ull;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        if (v[i] == value && (unsigned long long) i < best) best = (unsigned long long) i;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
        best = o < best ? o : best;
    }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(out_index, best);
}

// ---- process-wide cache of large device blocks --------------------------------------------------------------------
// One free list per device.  give(): a handle parks a block together with an event recorded on its stream after the
// block's last use.  take(): the smallest parked block that fits and is at most 25 % larger; the taker makes its stream
// wait on the event.  The cache holds at most kMaxBytes per device (DABRY_B200_CACHE_MB overrides; 0 disables).
class DeviceBlockCache {
public:
    static constexpr size_t kMinBytes = (size_t) 1 << 20;
    static DeviceBlockCache& instance() {
        static DeviceBlockCache cache;
        return cache;
    }
    void* take(int device, size_t bytes, size_t* got, cudaEvent_t* ready) {
        std::lock_guard<std::mutex> lock(mutex_);
        std::vector<Block>& v = blocks_[device];
        int best = -1;
        for (int i = 0; i < (int) v.size(); ++i)
            if (v[i].bytes >= bytes && v[i].bytes <= bytes + bytes / 4 && (best < 0 || v[i].bytes < v[best].bytes)) best = i;
        if (best < 0) return nullptr;
        const Block b = v[best];
        v.erase(v.begin() + best);
        held_[device] -= b.bytes;
        *got = b.bytes;
        *ready = b.ready;
        return b.p;
    }
    bool give(int device, void* p, size_t bytes, cudaEvent_t ready) {
        std::lock_guard<std::mutex> lock(mutex_);
        if (held_[device] + bytes > limit_) return false;
        blocks_[device].push_back(Block{p, bytes, ready});
        held_[device] += bytes;
        return true;
    }
    void trim(int device, cudaStream_t stream) {  // hand every parked block of the device back to the pool
        std::lock_guard<std::mutex> lock(mutex_);
        for (const Block& b : blocks_[device]) {
            cudaStreamWaitEvent(stream, b.ready, 0);
            cudaEventDestroy(b.ready);
            cudaFreeAsync(b.p, stream);
        }
        blocks_[device].clear();
        held_[device] = 0;
    }

private:
    struct Block {
        void* p;
        size_t bytes;
        cudaEvent_t ready;
    };
    DeviceBlockCache() {
        limit_ = (size_t) 32 << 30;  // per device; DABRY_B200_CACHE_MB overrides (a 10 M-extremal store is 61 GB)
        if (const char* env = getenv("DABRY_B200_CACHE_MB")) limit_ = (size_t) strtoull(env, nullptr, 10) << 20;
    }
    std::mutex mutex_;
    std::map<int, std::vector<Block>> blocks_;
    std::map<int, size_t> held_;
    size_t limit_;
};

// ---- process-wide pool of streams and events ----------------------------------------------------------------------
// The usual pattern is one handle per solve() call: creating and destroying two streams and a dozen events per solve is
// 0.4 ms when all goes well and now and then 3 - 23 ms of driver housekeeping (profiles/r02_e2e_outliers.txt).  A closed
// handle parks its (synchronised) streams and events here, the next handle on the same device takes them.  Parked sets
// are never destroyed: at process exit the driver reclaims them, and a static destructor must not call into a runtime
// that may already be gone.
struct StreamSet {
    static constexpr int kCopySlots = 8;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t copy_ev[kCopySlots] = {};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
};

class StreamPool {
public:
    static StreamPool& instance() {
        static StreamPool* pool = new StreamPool;  // leaked on purpose (see above)
        return *pool;
    }
    bool take(int device, StreamSet& out) {
        std::lock_guard<std::mutex> lock(mutex_);
        std::vector<StreamSet>& v = sets_[device];
        if (v.empty()) return false;
        out = std::move(v.back());
        v.pop_back();
        return true;
    }
    void give(int device, StreamSet&& set) {
        std::lock_guard<std::mutex> lock(mutex_);
        sets_[device].push_back(std::move(set));
    }

private:
    std::mutex mutex_;
    std::map<int, std::vector<StreamSet>> sets_;
};

// ---- backend ------------------------------------------------------------------------------------------------
class CudaBackend {
public:
    static const char* name() { return "cuda-sm100a"; }

    ~CudaBackend() {
        if (stream_) {
            // DABRY_B200_DEBUG_TIMING=1: host wall time of the parts of the destructor on stderr (diagnostic)
            const bool timing = getenv("DABRY_B200_DEBUG_TIMING") != nullptr;
            const auto t0 = std::chrono::steady_clock::now();
            cudaStreamSynchronize(stream_);
            if (copy_stream_) cudaStreamSynchronize(copy_stream_);
            const auto t1 = std::chrono::steady_clock::now();
            if (d_small_) cudaFreeAsync(d_small_, stream_);
            if (d_block_sum_) cudaFreeAsync(d_block_sum_, stream_);
            cudaStreamSynchronize(stream_);
            const auto t2 = std::chrono::steady_clock::now();
            // streams and events go back to the process-wide pool (idle: both streams were just synchronised)
            StreamSet set;
            set.stream = stream_;
            set.copy_stream = copy_stream_;
            set.ev[0] = ev0_;
            set.ev[1] = ev1_;
            set.ev[2] = tot0_;
            set.ev[3] = tot1_;
            for (int i = 0; i < kCopySlots; ++i) set.copy_ev[i] = copy_ev_[i];
            set.spans = std::move(spans_);
            StreamPool::instance().give(device_, std::move(set));
            if (timing) {
                const auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
                fprintf(stderr, "dabry backend close: sync %.2f ms, frees + sync %.2f ms, park %.2f ms\n", ms(t0, t1), ms(t1, t2),
                        ms(t2, std::chrono::steady_clock::now()));
            }
        }
    }

    // dabry_release_cached_memory: parked blocks of one device back to the driver
    static bool release_cached(int device) {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
            cudaGetLastError();
            return false;
        }
        int before = 0;
        cudaGetDevice(&before);
        cudaSetDevice(device);
        DeviceBlockCache::instance().trim(device, nullptr);
        cudaStreamSynchronize(nullptr);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        cudaSetDevice(before);
        return true;
    }

    // cudaHostRegister / cudaHostUnregister of a caller-owned array (dabry_host_register): portable, so that the
    // pinning holds for every device of a multi-GPU process
    static bool host_register(void* p, size_t bytes) {
        const cudaError_t rc = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
        if (rc == cudaErrorHostMemoryAlreadyRegistered) {
            cudaGetLastError();
            return true;
        }
        if (rc != cudaSuccess) cudaGetLastError();
        return rc == cudaSuccess;
    }
    static bool host_unregister(void* p) {
        const cudaError_t rc = cudaHostUnregister(p);
        if (rc != cudaSuccess) cudaGetLastError();
        return rc == cudaSuccess;
    }

    bool open(int device) {
        int count = 0;
        cudaError_t rc = cudaGetDeviceCount(&count);
        if (rc != cudaSuccess || count <= 0) {
            err_ = std::string("no CUDA device: ") + cudaGetErrorString(rc);
            cudaGetLastError();
            return false;
        }
        if (device < 0 || device >= count) {
            err_ = "CUDA device ordinal out of range";
            return false;
        }
        // two attribute reads instead of cudaGetDeviceProperties, which walks the whole property table (1 - 90 ms
        // measured per call on a busy host; a handle is opened for every solve)
        int major = 0, sms = 0;
        if (!chk(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device), "cudaDeviceGetAttribute") ||
            !chk(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device), "cudaDeviceGetAttribute"))
            return false;
        if (major != 10) {
            err_ = "device " + std::to_string(device) + " is compute capability " + std::to_string(major) +
                   ".x, not sm_100: this library ships sm_100a code only";
            return false;
        }
        device_ = device;
        sm_count_ = sms;
        if (!chk(cudaSetDevice(device), "cudaSetDevice")) return false;
        StreamSet set;
        if (StreamPool::instance().take(device, set)) {  // streams and events of a closed handle on this device
            stream_ = set.stream;
            copy_stream_ = set.copy_stream;
            ev0_ = set.ev[0];
            ev1_ = set.ev[1];
            tot0_ = set.ev[2];
            tot1_ = set.ev[3];
            for (int i = 0; i < kCopySlots; ++i) copy_ev_[i] = set.copy_ev[i];
            spans_ = std::move(set.spans);
            span_used_ = 0;
        } else {
            if (!chk(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
            cudaEventCreate(&ev0_);
            cudaEventCreate(&ev1_);
            cudaEventCreate(&tot0_);
            cudaEventCreate(&tot1_);
        }
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        if (!chk(cudaMallocAsync(&d_small_, 64, stream_), "cudaMallocAsync")) return false;
        cudaGetLastError();
        return true;
    }

    void activate() const {  // make this handle's device current (no-op when it already is)
        int current = -1;
        if (stream_ && (cudaGetDevice(&current) != cudaSuccess || current != device_)) cudaSetDevice(device_);
    }
    bool ok() const { return err_.empty(); }
    const std::string& error() const { return err_; }
    uint64_t launches() const { return launches_; }
    cudaStream_t stream() const { return stream_; }

    // Device memory.  Fresh blocks come stream-ordered from the device's default memory pool (unbounded release
    // threshold).  Blocks of 1 MB and more are NOT given back when a solver lets go of them: they go to a process-wide
    // per-device cache (DeviceBlockCache below) and the next solver of the same shape - the usual pattern is one handle
    // per solve() call - takes them from there.  Measured without it: dabry_create / dabry_destroy cost 3 - 40 ms per
    // solve with outliers of 170 - 420 ms when the driver trims and refills the pool (10 GB of site store).
    void* alloc(size_t bytes) {
        if (bytes == 0) bytes = 8;
        void* p = nullptr;
        if (bytes >= DeviceBlockCache::kMinBytes) {
            cudaEvent_t ready = nullptr;
            size_t got = 0;
            p = DeviceBlockCache::instance().take(device_, bytes, &got, &ready);
            if (p) {
                if (ready) {  // the block's last use was ordered on another handle's stream
                    cudaStreamWaitEvent(stream_, ready, 0);
                    cudaEventDestroy(ready);
                }
                live_[p] = got;
                return p;
            }
        }
        if (cudaMallocAsync(&p, bytes, stream_) != cudaSuccess) {
            cudaGetLastError();
            // out of memory: give the cached blocks of this device back and try once more
            DeviceBlockCache::instance().trim(device_, stream_);
            cudaStreamSynchronize(stream_);
            if (cudaMallocAsync(&p, bytes, stream_) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
        }
        if (bytes >= DeviceBlockCache::kMinBytes) live_[p] = bytes;
        return p;
    }
    template <class T>
    void free(T* ptr) {
        void* p = (void*) ptr;
        if (!p) return;
        auto it = live_.find(p);
        if (it == live_.end()) {
            cudaFreeAsync(p, stream_);
            return;
        }
        const size_t bytes = it->second;
        live_.erase(it);
        cudaEvent_t ready = nullptr;
        if (cudaEventCreateWithFlags(&ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(ready, stream_) != cudaSuccess) {
            cudaGetLastError();
            if (ready) cudaEventDestroy(ready);
            cudaFreeAsync(p, stream_);
            return;
        }
        if (!DeviceBlockCache::instance().give(device_, p, bytes, ready)) {
            cudaEventDestroy(ready);
            cudaFreeAsync(p, stream_);
        }
    }
    void upload(void* d, const void* h, size_t bytes) {
        chk(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream_), "H2D copy");
        chk(cudaStreamSynchronize(stream_), "H2D sync");
    }
    // background H2D copy of one chunk on the copy stream; wait_chunk orders the compute stream behind it
    static constexpr int kCopySlots = StreamSet::kCopySlots;
    void upload_chunk_async(void* d, const void* h, size_t bytes, int slot) {
        if (!copy_stream_) {
            if (!chk(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking), "cudaStreamCreate")) return;
            for (int i = 0; i < kCopySlots; ++i) cudaEventCreateWithFlags(&copy_ev_[i], cudaEventDisableTiming);
        }
        // the staging buffer was allocated on the compute stream: the copy may not start before that allocation
        cudaEventRecord(copy_ev_[slot], stream_);
        cudaStreamWaitEvent(copy_stream_, copy_ev_[slot], 0);
        chk(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, copy_stream_), "H2D copy (streamed)");
        chk(cudaEventRecord(copy_ev_[slot], copy_stream_), "event record");
    }
    void wait_chunk(int slot) {
        if (copy_stream_) chk(cudaStreamWaitEvent(stream_, copy_ev_[slot], 0), "stream wait");
    }
    void sync_copies() {
        if (copy_stream_) cudaStreamSynchronize(copy_stream_);
    }
    static bool host_is_pinned(const void* p) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return attr.type == cudaMemoryTypeHost;
    }
    void download(void* h, const void* d, size_t bytes) {
        chk(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream_), "D2H copy");
        chk(cudaStreamSynchronize(stream_), "D2H sync");
    }
    void zero(void* d, size_t bytes) { chk(cudaMemsetAsync(d, 0, bytes, stream_), "memset"); }
    void fill_ones(void* d, size_t bytes) { chk(cudaMemsetAsync(d, 0xFF, bytes, stream_), "memset"); }
    void copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows) {
        chk(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToDevice, stream_), "D2D copy");
    }
    void sync() { chk(cudaStreamSynchronize(stream_), "stream sync"); }
    // the problem description of the calling engine -> constant memory, ordered on the engine's stream
    void set_problem(const ProblemDev& P) {
        chk(cudaMemcpyToSymbolAsync(c_problem, &P, sizeof(ProblemDev), 0, cudaMemcpyHostToDevice, stream_), "bind problem");
    }

    void upload_async(void* d, const void* h, size_t bytes) {
        chk(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream_), "H2D copy");
    }

    // ---- process-wide state of one device: the problem description of every handle goes through ONE __constant__
    // symbol (model.cuh), so the entry points of all handles on a device are serialised (CallGuard, taken by every
    // C entry point), and a handle that binds its problem first makes its stream wait for the kernels of the handle
    // that used the constants before it.  Handles on different devices run concurrently.
    struct DeviceState {
        std::mutex mutex;
        cudaEvent_t last_use = nullptr;
        const void* last_owner = nullptr;
    };
    static DeviceState& device_state(int device) {
        static std::mutex table_mutex;
        static std::map<int, DeviceState> table;
        std::lock_guard<std::mutex> lock(table_mutex);
        return table[device];
    }
    class CallGuard {
    public:
        explicit CallGuard(CudaBackend& be) : be_(be), st_(device_state(be.device_)) {
            st_.mutex.lock();
            be_.activate();
            if (be_.stream_ && st_.last_owner && st_.last_owner != &be_ && st_.last_use)
                cudaStreamWaitEvent(be_.stream_, st_.last_use, 0);
        }
        ~CallGuard() {
            if (be_.stream_) {
                if (!st_.last_use) cudaEventCreateWithFlags(&st_.last_use, cudaEventDisableTiming);
                if (st_.last_use) cudaEventRecord(st_.last_use, be_.stream_);
                st_.last_owner = &be_;
            }
            st_.mutex.unlock();
        }
        CallGuard(const CallGuard&) = delete;
        CallGuard& operator=(const CallGuard&) = delete;

    private:
        CudaBackend& be_;
        DeviceState& st_;
    };

    template <class Fn>
    static void destroy_locked(CudaBackend& be, Fn&& destroy) {  // dabry_destroy: the backend dies inside the lock
        DeviceState& st = device_state(be.device_);
        std::lock_guard<std::mutex> lock(st.mutex);
        be.activate();
        if (st.last_owner == &be) st.last_owner = nullptr;  // ~CudaBackend drains its stream
        destroy();
    }

    // ---- collectives: NCCL on the solver's stream (nccl_dyn.cuh) ---------------------------------------------------
    struct Comm {
        nccl::Comm_t comm = nullptr;
        int rank = 0, world = 1, device = 0;
    };
    static int comm_rank(const Comm* c) { return c->rank; }
    static int comm_world(const Comm* c) { return c->world; }
    static bool comm_unique_id(void* id, std::string& err) {
        const nccl::Api& n = nccl::api();
        if (!n.ok()) {
            err = n.error;
            return false;
        }
        nccl::UniqueId uid;
        const int rc = n.GetUniqueId(&uid);
        if (rc != 0) {
            err = std::string("ncclGetUniqueId: ") + n.GetErrorString(rc);
            return false;
        }
        memcpy(id, &uid, sizeof(uid));
        return true;
    }
    static Comm* comm_create(int device, int rank, int world, const void* id, std::string& err) {
        const nccl::Api& n = nccl::api();
        if (!n.ok()) {
            err = n.error;
            return nullptr;
        }
        if (cudaSetDevice(device) != cudaSuccess) {
            cudaGetLastError();
            err = "dabry_comm_create: bad device";
            return nullptr;
        }
        nccl::UniqueId uid;
        memcpy(&uid, id, sizeof(uid));
        Comm* c = new Comm;
        c->rank = rank;
        c->world = world;
        c->device = device;
        const int rc = n.CommInitRank(&c->comm, world, uid, rank);
        if (rc != 0) {
            err = std::string("ncclCommInitRank: ") + n.GetErrorString(rc);
            delete c;
            return nullptr;
        }
        return c;
    }
    static void comm_destroy(Comm* c) {
        if (!c) return;
        if (c->comm) {
            cudaSetDevice(c->device);
            nccl::api().CommDestroy(c->comm);
        }
        delete c;
    }
    bool nccl_ok(int rc, const char* what) {
        if (rc == 0) return true;
        if (err_.empty()) err_ = std::string(what) + ": " + nccl::api().GetErrorString(rc);
        return false;
    }
    // ring: n doubles to rank - 1, n doubles from rank + 1, one group (ncclGroupStart / End) so that neither blocks
    void comm_ring_shift(Comm* c, const double* send, double* recv, int64_t n) {
        const nccl::Api& a = nccl::api();
        const int to = (c->rank + c->world - 1) % c->world, from = (c->rank + 1) % c->world;
        nccl_ok(a.GroupStart(), "ncclGroupStart");
        nccl_ok(a.Send(send, (size_t) n, nccl::kFloat64, to, c->comm, stream_), "ncclSend");
        nccl_ok(a.Recv(recv, (size_t) n, nccl::kFloat64, from, c->comm, stream_), "ncclRecv");
        nccl_ok(a.GroupEnd(), "ncclGroupEnd");
    }
    void comm_allreduce_min_f64(Comm* c, double* buf, int64_t n) {
        nccl_ok(nccl::api().AllReduce(buf, buf, (size_t) n, nccl::kFloat64, nccl::kMin, c->comm, stream_), "ncclAllReduce");
    }
    // recv holds world x bytes; in place when send == recv + rank * bytes
    void comm_allgather(Comm* c, const void* send, void* recv, size_t bytes) {
        nccl_ok(nccl::api().AllGather(send, recv, bytes, nccl::kUint8, c->comm, stream_), "ncclAllGather");
    }

    // device time of stretches that must not wait for the device (the collectives): event pairs, read at span_collect
    void span_begin() {
        if (span_used_ == spans_.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            spans_.push_back({a, b});
        }
        cudaEventRecord(spans_[span_used_].first, stream_);
    }
    void span_end() { cudaEventRecord(spans_[span_used_++].second, stream_); }
    double span_collect() {
        double total = 0.;
        for (size_t i = 0; i < span_used_; ++i) {
            float ms = 0.f;
            if (cudaEventSynchronize(spans_[i].second) == cudaSuccess &&
                cudaEventElapsedTime(&ms, spans_[i].first, spans_[i].second) == cudaSuccess)
                total += ms;
            else
                cudaGetLastError();
        }
        span_used_ = 0;
        return total;
    }

    template <class F>
    void launch(int64_t n, const F& f) {
        if (n <= 0 || !ok()) return;
        static_assert(sizeof(F) <= 4000, "phase functor must fit the kernel parameter space");
        const int64_t blocks = (n + F::kThreads - 1) / F::kThreads;
        phase_kernel<F><<<(unsigned) blocks, F::kThreads, 0, stream_>>>(n, f);
        ++launches_;
        chk(cudaGetLastError(), "kernel launch");
    }

    // exclusive prefix sum of (request >= 0) into offset; optional dense list of the flagged indices; returns total
    int64_t scan_requests(const int32_t* request, int64_t n, int32_t* offset, int32_t* list,
                          const int32_t* values = nullptr, int32_t value_base = 0) {
        if (n <= 0 || !ok()) return 0;
        const int64_t blocks = (n + kScanBlock - 1) / kScanBlock;
        if (blocks > block_sum_cap_) {
            if (d_block_sum_) cudaFreeAsync(d_block_sum_, stream_);
            block_sum_cap_ = blocks + blocks / 2 + 64;
            if (!chk(cudaMallocAsync(&d_block_sum_, sizeof(int32_t) * block_sum_cap_, stream_), "cudaMallocAsync")) return 0;
        }
        scan_pass1<<<(unsigned) blocks, kScanBlock, 0, stream_>>>(request, n, offset, d_block_sum_);
        scan_pass2<<<1, kScanBlock, 0, stream_>>>(d_block_sum_, blocks, (int64_t*) d_small_);
        scan_pass3<<<(unsigned) blocks, kScanBlock, 0, stream_>>>(request, n, offset, d_block_sum_, list, values, value_base);
        launches_ += 3;
        chk(cudaGetLastError(), "scan launch");
        int64_t total = 0;
        download(&total, d_small_, sizeof(total));
        return total;
    }

    double min_f64(const double* v, int64_t n) {
        if (n <= 0 || !ok()) return INFINITY;
        unsigned long long key = ~0ull;
        upload(d_small_, &key, sizeof(key));
        const int blocks = (int) (n < 256ll * 4 * sm_count_ ? (n + 255) / 256 : 4 * sm_count_);
        min_f64_kernel<<<blocks, 256, 0, stream_>>>(v, n, (unsigned long long*) d_small_);
        ++launches_;
        chk(cudaGetLastError(), "min launch");
        download(&key, d_small_, sizeof(key));
        const unsigned long long b = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;
        double out;
        memcpy(&out, &b, sizeof(out));
        return out;
    }

    int64_t first_equal_f64(const double* v, int64_t n, double value) {
        if (n <= 0 || !ok()) return -1;
        unsigned long long idx = ~0ull;
        upload(d_small_, &idx, sizeof(idx));
        const int blocks = (int) (n < 256ll * 4 * sm_count_ ? (n + 255) / 256 : 4 * sm_count_);
        first_equal_kernel<<<blocks, 256, 0, stream_>>>(v, n, value, (unsigned long long*) d_small_);
        ++launches_;
        chk(cudaGetLastError(), "argmin launch");
        download(&idx, d_small_, sizeof(idx));
        return idx == ~0ull ? -1 : (int64_t) idx;
    }

    // device time of a phase: CUDA events on the launching stream
    void tic() { cudaEventRecord(ev0_, stream_); }
    double toc() {
        cudaEventRecord(ev1_, stream_);
        if (!chk(cudaEventSynchronize(ev1_), "phase sync")) return 0.;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0_, ev1_);
        return ms;
    }
    void tic_total() { cudaEventRecord(tot0_, stream_); }
    double toc_total() {
        cudaEventRecord(tot1_, stream_);
        if (!chk(cudaEventSynchronize(tot1_), "solve sync")) return 0.;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, tot0_, tot1_);
        return ms;
    }

private:
    bool chk(cudaError_t rc, const char* what) {
        if (rc == cudaSuccess) return true;
        if (err_.empty()) err_ = std::string(what) + ": " + cudaGetErrorString(rc);
        return false;
    }
    int device_ = 0, sm_count_ = 148;
    cudaStream_t stream_ = nullptr, copy_stream_ = nullptr;
    cudaEvent_t copy_ev_[kCopySlots] = {};
    std::unordered_map<void*, size_t> live_;  // cacheable blocks this handle holds -> their real size
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, tot0_ = nullptr, tot1_ = nullptr;
    void* d_small_ = nullptr;
    int32_t* d_block_sum_ = nullptr;
    int64_t block_sum_cap_ = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans_;
    size_t span_used_ = 0;
    uint64_t launches_ = 0;
    std::string err_;
};

}  // namespace dabry

#define DABRY_BACKEND dabry::CudaBackend
#include "capi.inl"
