// libtoeplitz_b200.so: handles, plans and the extern "C" entry points of
// include/toeplitz_b200.h.  Host-side orchestration only -- all arithmetic is in the
// kernels of fft_kernels.cuh / aux_kernels.cuh / blas1.cuh / levinson.cu.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <type_traits>
#include <vector>

#include <cuda.h>

#include "capi_internal.h"
#include "fft_launch.cuh"
#include "aux_kernels.cuh"
#include "blas1.cuh"
#include "krylov_batched.cuh"

namespace tb {

// ---------------------------------------------------------------------------------------
// errors, counters, options
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<long long> g_group_bytes{48ll << 20};   // two-level scratch budget per group
static std::atomic<long long> g_host_chunk_bytes{256ll << 20};
static std::atomic<long long> g_krylov_work_bytes{16ll << 30};   // work-vector budget of the batched cgs / cg (columns per chunk)
static std::atomic<int> g_time_passes{0};
static std::atomic<int> g_ca_fuse{0};            // 1: pass C of a group + pass A of the next one on its stream as one launch; 2: + async x tile
static std::atomic<int> g_pdl{1};                // programmatic dependent launch of passes B / C: 0 off, 1 products without group overlap, 2 always
static std::atomic<int> g_l2_persist_mb{64};     // persisting-L2 carve-out for the two-level workspaces (0 = off); +4 % on config 2
static std::atomic<int> g_rows_direct{1};        // pass B addresses its workspace rows directly (IO_SPEC path) instead of through the column views
static std::atomic<int> g_pair_columns{1};       // real columns two per complex transform (0: one per transform, columns numerically independent)
static std::atomic<int> g_tma{0};                // TMA-staged strided passes (tma_cols.cuh) where eligible; measured, see DESIGN.md
extern int g_lev_gen;                            // levinson.cu
int g_prune_half = 1;                            // pruned pass A / pass C when half of the embedding is padding
static std::atomic<int> g_overlap_streams{2};   // helper streams for group overlap (0/1 = off)
static std::atomic<int> g_l2max_f64{12}, g_l2max_f32{13}, g_ta_cap{0}, g_force_l1{0};
static std::atomic<int> g_single_max_f64{13}, g_single_max_f32{14};   // longest embedding done as ONE shared-memory transform (2^k)   // plan tuning knobs

// optional per-pass device timing (bench.py's roofline leg): CUDA events recorded on the
// launching stream around every pass while "time_passes" is on.
struct PassEvents { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev; };
static std::mutex g_pass_mutex;
static PassEvents g_pass[5];      // 0 = cols A, 1 = rows B, 2 = cols C, 3 = single-kernel rows, 4 = fused cols C+A

struct PassTimer {
    int which; cudaStream_t s; cudaEvent_t a = nullptr, b = nullptr; bool on;
    PassTimer(int w, cudaStream_t st) : which(w), s(st), on(g_time_passes.load() != 0) {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
    }
    ~PassTimer() {
        if (on) {
            cudaEventRecord(b, s);
            std::lock_guard<std::mutex> lock(g_pass_mutex);
            g_pass[which].ev.emplace_back(a, b);
        }
    }
};

int set_error(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}
int cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return TB200_OK;
    if (e == cudaErrorMemoryAllocation) return set_error(TB200_OOM, "%s: %s", what, cudaGetErrorString(e));
    return set_error(TB200_CUDA_ERR, "%s: %s", what, cudaGetErrorString(e));
}
void count_launch(int n) { g_launches += n; }

#define TB_CUDA(call, what)                                   \
    do {                                                      \
        int st__ = cuda_status((call), what);                 \
        if (st__ != TB200_OK) return st__;                    \
    } while (0)
#define TB_TRY(expr)                                          \
    do {                                                      \
        int st__ = (expr);                                    \
        if (st__ != TB200_OK) return st__;                    \
    } while (0)

static inline bool dt_is_cplx(int dt) { return dt == TB200_C32 || dt == TB200_C64; }
static inline int dt_prec(int dt) { return (dt == TB200_F64 || dt == TB200_C64) ? 1 : 0; }
static inline size_t dt_size(int dt) { return (dt_prec(dt) ? 8 : 4) * (dt_is_cplx(dt) ? 2 : 1); }
static inline bool dt_valid(int dt) { return dt >= 0 && dt <= 3; }

// Device memory comes from the stream-ordered allocator on one private stream per device, with the
// pool's release threshold raised so that factorize/destroy cycles (the reference re-factorizes on
// every vector mul!, src/linearalgebra.jl:37) reuse cached blocks instead of paying cudaMalloc/cudaFree.
static cudaStream_t alloc_stream() {
    static std::mutex mu;
    static cudaStream_t streams[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!streams[dev]) {
        cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    return streams[dev];
}
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t st = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
};
// Ordering rules of the pool (one private allocation stream per device):
//  * dev_alloc_on(.., s): the block is allocated on the private stream and stream `s` is made to wait for that
//    allocation (event), so kernels launched on `s` afterwards may use it -- no host synchronisation.
//  * dev_alloc_sync(..): host-synchronous variant for data that is cached per device and read from ANY stream later
//    (twiddle tables): the block and its contents are complete when the call returns.
//  * frees (DevBuf destructor) run on the private stream; before a handle or a workspace is released,
//    retire_after(s) makes the private stream wait for everything already queued on each stream that used it, so a
//    block can never be handed out again while a kernel of the old owner is still in flight (tb200_destroy right after
//    an asynchronous tb200_mul is safe on any stream).
static int order_stream_after(cudaStream_t waiter, cudaStream_t producer, const char* what) {
    cudaEvent_t ev = nullptr;
    TB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), what);
    cudaError_t e = cudaEventRecord(ev, producer);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(waiter, ev, 0);
    cudaEventDestroy(ev);                       // released by the runtime once the record has completed
    return cuda_status(e, what);
}
static int retire_after(cudaStream_t user) { return order_stream_after(alloc_stream(), user, "retire ordering"); }
static int dev_alloc_raw(std::shared_ptr<DevBuf>& out, size_t bytes, const char* what) {
    auto b = std::make_shared<DevBuf>();
    if (bytes == 0) bytes = 16;
    b->st = alloc_stream();
    cudaError_t e = cudaMallocAsync(&b->p, bytes, b->st);
    if (e != cudaSuccess) { b->p = nullptr; return cuda_status(e, what); }
    b->bytes = bytes;
    out = b;
    return TB200_OK;
}
static int dev_alloc_on(std::shared_ptr<DevBuf>& out, size_t bytes, const char* what, cudaStream_t s) {
    TB_TRY(dev_alloc_raw(out, bytes, what));
    return order_stream_after(s, alloc_stream(), what);
}
static int dev_alloc_sync(std::shared_ptr<DevBuf>& out, size_t bytes, const char* what) {
    TB_TRY(dev_alloc_raw(out, bytes, what));
    return cuda_status(cudaStreamSynchronize(alloc_stream()), what);
}
// host -> device upload that is complete (visible to every stream) when the call returns
static int upload_sync(void* dst, const void* src, size_t bytes, const char* what) {
    TB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, alloc_stream()), what);
    return cuda_status(cudaStreamSynchronize(alloc_stream()), what);
}

static int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}
static inline unsigned ew_grid(long long n, int threads = 256) {
    long long g = (n + threads - 1) / threads;
    long long cap = (long long)num_sms() * 16;
    return (unsigned)std::max(1ll, std::min(g, cap));
}

// ---------------------------------------------------------------------------------------
// twiddle tables
// ---------------------------------------------------------------------------------------
template <typename R>
static void fill_twiddles(std::vector<cx<R>>& v, long long count, long long stride, long long period) {
    // v[k] = exp(-2 pi i (k*stride)/period), evaluated in long double on an exactly reduced argument
    v.resize((size_t)count);
    const long double twopi = 6.283185307179586476925286766559005768L;
    for (long long k = 0; k < count; ++k) {
        long long num = (k * stride) % period;
        // reduce to the first octant for accuracy
        long double frac = (long double)num / (long double)period;   // exact: period is a power of two
        long double c = cosl(twopi * frac), s = sinl(twopi * frac);
        // exact values on the axes / diagonals
        if (num * 4 % period == 0) {
            long long q = num * 4 / period;
            c = (q == 0) ? 1 : (q == 2) ? -1 : 0;
            s = (q == 1) ? 1 : (q == 3) ? -1 : 0;
        }
        v[(size_t)k] = mk<R>((R)c, (R)(-s));
    }
}

struct TwTables {
    std::shared_ptr<DevBuf> f32[MAX_LOG2L + 1], f64[MAX_LOG2L + 1];
};
static std::mutex g_tw_mutex;
static TwTables g_tw[64];

// stage twiddle table (thread order, see fft_core.cuh) for an L = 2^log2l transform
template <typename R>
static int stage_twiddles(int log2l, const cx<R>** out) {
    int dev = 0;
    TB_CUDA(cudaGetDevice(&dev), "cudaGetDevice");
    constexpr int LOGE = le_of<R>(), E = 1 << LOGE;
    if (log2l < LOGE || log2l > MAX_LOG2L) return set_error(TB200_UNSUPPORTED, "no stage table for 2^%d", log2l);
    std::lock_guard<std::mutex> lock(g_tw_mutex);
    std::shared_ptr<DevBuf>& slot = (sizeof(R) == 8) ? g_tw[dev].f64[log2l] : g_tw[dev].f32[log2l];
    if (!slot) {
        std::vector<cx<R>> h((size_t)std::max(stage_tw_total(log2l, LOGE), 1));
        const long double twopi = 6.283185307179586476925286766559005768L;
        for (int k = 0; k < log2l / LOGE; ++k) {
            const int ls = log2l - LOGE * (k + 1);
            if (ls <= 0) continue;
            const long long s = 1ll << ls, period = s * E;
            const int off = stage_tw_off(log2l, k, LOGE);
            for (int c = 1; c < E; ++c)
                for (long long r = 0; r < s; ++r) {
                    const long long num = (r * c) % period;
                    long double cs = cosl(twopi * (long double)num / (long double)period);
                    long double sn = sinl(twopi * (long double)num / (long double)period);
                    if (num * 4 % period == 0) {
                        const long long q = num * 4 / period;
                        cs = (q == 0) ? 1 : (q == 2) ? -1 : 0;
                        sn = (q == 1) ? 1 : (q == 3) ? -1 : 0;
                    }
                    h[(size_t)off + (size_t)(c - 1) * s + r] = mk<R>((R)cs, (R)(-sn));
                }
        }
        std::shared_ptr<DevBuf> b;
        TB_TRY(dev_alloc_sync(b, sizeof(cx<R>) * h.size(), "twiddle table"));
        TB_TRY(upload_sync(b->p, h.data(), sizeof(cx<R>) * h.size(), "twiddle upload"));
        slot = b;
    }
    *out = reinterpret_cast<const cx<R>*>(slot->p);
    return TB200_OK;
}

}  // namespace tb

using namespace tb;

// ---------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------
// Everything a product needs besides the (read-only) spectrum and twiddle tables, owned PER CALLING STREAM: the
// two-level workspaces, the helper streams that overlap consecutive column groups, and the Bluestein buffers.  One
// handle can therefore be used from several streams / host threads at once -- the reference's factorization cannot
// (its single `tmp` vector is shared by every product, src/ToeplitzMatrices.jl:97-101).
struct StreamWork {
    static constexpr int MAXS = 4;
    std::shared_ptr<DevBuf> scratch, scratch_extra[MAXS], bs_u, bs_v;
    cudaStream_t hs[MAXS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[MAXS] = {nullptr, nullptr, nullptr, nullptr};
    bool win_set[MAXS] = {false, false, false, false};     // a persisting-L2 window is currently set on hs[i] (or the user stream)
    ~StreamWork() {
        for (int i = 0; i < MAXS; ++i) {
            if (hs[i]) cudaStreamDestroy(hs[i]);
            if (ev_join[i]) cudaEventDestroy(ev_join[i]);
        }
        if (ev_fork) cudaEventDestroy(ev_fork);
    }
};

struct tb200_fact {
    int kind = 0, dtype = 0, prec = 0, a_real = 0, device = 0;
    int64_t m = 0, n = 0, Np = 0;
    int log2np = 0, two_level = 0, log2l1 = 0, log2l2 = 0, ta = 0;
    // three-level plan (embeddings beyond the two-level limit): an outer strided pass of 2^log2l1 points around a
    // batched two-level transform of the 2^log2l2-point rows, which lives in the child handle `inner`
    int three_level = 0;
    std::shared_ptr<tb200_fact> inner;
    int reverse_x = 0;
    static constexpr int MAXS = StreamWork::MAXS;
    std::shared_ptr<DevBuf> spec, inv_spec, tw_lo, tw_hi, tw_g;
    int tw_hb = 0;
    bool spec_valid = false;
    // readiness of the spectrum / reciprocal spectrum: recorded on the stream that produced them; a product on any
    // OTHER stream waits for the event first (tb200_mul_host's internal streams included)
    cudaEvent_t ev_spec = nullptr, ev_inv = nullptr;
    std::atomic<bool> spec_done{false}, inv_done{false};      // the event has been seen complete: no more waits
    cudaStream_t spec_stream = nullptr, inv_stream = nullptr;
    // Bluestein plan (Circulant, n not a power of two): spectrum kept in NATURAL order (n entries)
    bool bluestein = false;
    std::shared_ptr<tb200_fact> bs_conv;          // complex symmetric Toeplitz T[d] = exp(+i pi d^2/n)
    std::shared_ptr<DevBuf> bs_chirp;             // chirp c[j] = exp(-i pi j^2/n)
    cudaStream_t host_s[3] = {nullptr, nullptr, nullptr};   // tb200_mul_host: upload / compute / download streams
    std::mutex mu;                                // guards `works`, inv_spec creation
    std::vector<std::pair<cudaStream_t, std::unique_ptr<StreamWork>>> works;
    StreamWork* work_for(cudaStream_t s) {
        std::lock_guard<std::mutex> lock(mu);
        for (auto& w : works) if (w.first == s) return w.second.get();
        works.emplace_back(s, std::unique_ptr<StreamWork>(new StreamWork()));
        return works.back().second.get();
    }
    ~tb200_fact() {
        // frees are stream-ordered on the pool's private stream: let it wait for everything queued on the streams that
        // used this handle (helper streams always join their user stream before a call returns)
        for (auto& w : works) retire_after(w.first);
        if (ev_spec) { retire_after(spec_stream); cudaEventDestroy(ev_spec); }
        if (ev_inv) cudaEventDestroy(ev_inv);
        for (int i = 0; i < 3; ++i) if (host_s[i]) cudaStreamDestroy(host_s[i]);
    }
};

namespace tb {

// A handle is bound to the device it was created on; entry points switch to it for the call.
struct DeviceGuard {
    int prev = -1; bool switched = false;
    explicit DeviceGuard(const tb200_fact* h) {
        if (!h) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != h->device) { cudaSetDevice(h->device); switched = true; }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

static int next_log2(int64_t v) { int l = 0; while ((1ll << l) < v) ++l; return l; }

static int make_plan(tb200_fact* h) {
    const int l = h->log2np;
    const int max_single = h->prec ? g_single_max_f64.load() : g_single_max_f32.load();
    if (l < 4) return set_error(TB200_UNSUPPORTED, "embedding shorter than 16 is handled by tb200_mul_direct");
    if (l <= max_single) {
        h->two_level = 0; h->log2l1 = 0; h->log2l2 = l; h->ta = 0;
        return TB200_OK;
    }
    int l2max = h->prec ? g_l2max_f64.load() : g_l2max_f32.load();
    if (!h->prec && l >= 25 && l2max < 14) l2max = 14;        // 2^25: 2^11 x 2^14 measured 12 % faster than 2^12 x 2^13
    if (h->prec && l - l2max > 11 && l2max < 13) l2max = 13;  // keep 128-byte... at least 64-byte column tiles as long as possible
    int l1 = std::max(l - l2max, std::min(9, l / 2));
    int l2 = l - l1;
    const int l1max = 12;                                    // 2^12 x 2^13 (Float64, 2 columns per tile) / 2^12 x 2^14 (Float32)
    if (l1 > l1max) { l1 = l1max; l2 = l - l1; }
    if (l2 > (h->prec ? 13 : 14)) {
        // beyond the two-level plan: N' = 2^9 x 2^(l-9); the rows are themselves two-level transforms (child plan)
        const int lo = 9, lr = l - lo;
        if (l > 30 || lr > l1max + (h->prec ? 13 : 14))
            return set_error(TB200_UNSUPPORTED, "embedding length 2^%d exceeds the three-level plan (max 2^30)", l);
        h->three_level = 1; h->two_level = 1; h->log2l1 = lo; h->log2l2 = lr;
        h->ta = 128 / (int)(h->prec ? 16 : 8);
        std::shared_ptr<tb200_fact> child(new tb200_fact());
        child->kind = TB200_CIRCULANT; child->dtype = h->prec ? TB200_C64 : TB200_C32; child->prec = h->prec; child->a_real = 0;
        child->device = h->device; child->m = child->n = child->Np = 1ll << lr; child->log2np = lr;
        TB_TRY(make_plan(child.get()));
        h->inner = child;
        h->tw_hb = (l + 1) / 2;
        const long long nlo3 = 1ll << h->tw_hb, nhi3 = 1ll << (l - h->tw_hb);
        if (h->prec) {
            std::vector<cx<double>> lo_t, hi_t;
            fill_twiddles<double>(lo_t, nlo3, 1, 1ll << l);
            fill_twiddles<double>(hi_t, nhi3, nlo3, 1ll << l);
            TB_TRY(dev_alloc_sync(h->tw_lo, lo_t.size() * 16, "level twiddles"));
            TB_TRY(dev_alloc_sync(h->tw_hi, hi_t.size() * 16, "level twiddles"));
            TB_TRY(upload_sync(h->tw_lo->p, lo_t.data(), lo_t.size() * 16, "level twiddle upload"));
            TB_TRY(upload_sync(h->tw_hi->p, hi_t.data(), hi_t.size() * 16, "level twiddle upload"));
        } else {
            std::vector<cx<float>> lo_t, hi_t;
            fill_twiddles<float>(lo_t, nlo3, 1, 1ll << l);
            fill_twiddles<float>(hi_t, nhi3, nlo3, 1ll << l);
            TB_TRY(dev_alloc_sync(h->tw_lo, lo_t.size() * 8, "level twiddles"));
            TB_TRY(dev_alloc_sync(h->tw_hi, hi_t.size() * 8, "level twiddles"));
            TB_TRY(upload_sync(h->tw_lo->p, lo_t.data(), lo_t.size() * 8, "level twiddle upload"));
            TB_TRY(upload_sync(h->tw_hi->p, hi_t.data(), hi_t.size() * 8, "level twiddle upload"));
        }
        return TB200_OK;                               // no [E][N2] table at this N2: the passes use the split tables per element
    }
    if (l1 < 6) { l1 = 6; l2 = l - l1; }
    if (g_force_l1.load() >= 6 && l - g_force_l1.load() >= 4 && l - g_force_l1.load() <= (h->prec ? 13 : 14)) { l1 = g_force_l1.load(); l2 = l - l1; }
    const int full = 128 / (int)(h->prec ? 16 : 8);            // columns per 128-byte line
    int ta = full;
    const size_t esz = h->prec ? 16 : 8;
    while (ta > 2 && ((size_t)1 << l1) * ta * esz > 128 * 1024) ta >>= 1;
    if (g_ta_cap.load() >= 4 && ta > g_ta_cap.load()) ta = g_ta_cap.load();
    h->two_level = 1; h->log2l1 = l1; h->log2l2 = l2; h->ta = ta;
    // inter-level twiddles  w_Np^m = lo[m & mask] * hi[m >> hb]  (cached per device / precision / length)
    h->tw_hb = (l + 1) / 2;
    const long long nlo = 1ll << h->tw_hb, nhi = 1ll << (l - h->tw_hb);
    {
        static std::mutex mu;
        static std::shared_ptr<DevBuf> cache[64][2][40][2];
        std::lock_guard<std::mutex> lock(mu);
        auto& clo = cache[h->device & 63][h->prec][l][0];
        auto& chi = cache[h->device & 63][h->prec][l][1];
        if (!clo || !chi) {
            if (h->prec) {
                std::vector<cx<double>> lo, hi;
                fill_twiddles<double>(lo, nlo, 1, 1ll << l);
                fill_twiddles<double>(hi, nhi, nlo, 1ll << l);
                TB_TRY(dev_alloc_sync(clo, lo.size() * 16, "level twiddles"));
                TB_TRY(dev_alloc_sync(chi, hi.size() * 16, "level twiddles"));
                TB_TRY(upload_sync(clo->p, lo.data(), lo.size() * 16, "level twiddle upload"));
                TB_TRY(upload_sync(chi->p, hi.data(), hi.size() * 16, "level twiddle upload"));
            } else {
                std::vector<cx<float>> lo, hi;
                fill_twiddles<float>(lo, nlo, 1, 1ll << l);
                fill_twiddles<float>(hi, nhi, nlo, 1ll << l);
                TB_TRY(dev_alloc_sync(clo, lo.size() * 8, "level twiddles"));
                TB_TRY(dev_alloc_sync(chi, hi.size() * 8, "level twiddles"));
                TB_TRY(upload_sync(clo->p, lo.data(), lo.size() * 8, "level twiddle upload"));
                TB_TRY(upload_sync(chi->p, hi.data(), hi.size() * 8, "level twiddle upload"));
            }
        }
        h->tw_lo = clo;
        h->tw_hi = chi;
        // [E][N2] table w_Np^(delta_i * j2): delta_i = frequency offset of register i in the strided transform
        static std::shared_ptr<DevBuf> gcache[64][2][40][16];
        auto& cg = gcache[h->device & 63][h->prec][l][l1];
        if (!cg) {
            const int le = h->prec ? le_of<double>() : le_of<float>();
            const int Ecount = 1 << le;
            const long long N2 = 1ll << l2, Npl = 1ll << l;
            PlanDims d1{l1, le};
            std::vector<long long> nums((size_t)Ecount * N2);
            for (int i = 0; i < Ecount; ++i) {
                const long long delta = d1.freq_of_pos(d1.pos_of_reg(i, 0));
                for (long long j2 = 0; j2 < N2; ++j2) nums[(size_t)i * N2 + j2] = (delta * j2) % Npl;
            }
            const long double twopi = 6.283185307179586476925286766559005768L;
            if (h->prec) {
                std::vector<cx<double>> g(nums.size());
                for (size_t e = 0; e < nums.size(); ++e) {
                    const long double ang = twopi * (long double)nums[e] / (long double)Npl;
                    g[e] = mk<double>((double)cosl(ang), (double)(-sinl(ang)));
                }
                TB_TRY(dev_alloc_sync(cg, g.size() * 16, "level twiddle table"));
                TB_TRY(upload_sync(cg->p, g.data(), g.size() * 16, "level twiddle upload"));
            } else {
                std::vector<cx<float>> g(nums.size());
                for (size_t e = 0; e < nums.size(); ++e) {
                    const long double ang = twopi * (long double)nums[e] / (long double)Npl;
                    g[e] = mk<float>((float)cosl(ang), (float)(-sinl(ang)));
                }
                TB_TRY(dev_alloc_sync(cg, g.size() * 8, "level twiddle table"));
                TB_TRY(upload_sync(cg->p, g.data(), g.size() * 8, "level twiddle upload"));
            }
        }
        h->tw_g = cg;
    }
    return TB200_OK;
}

// Persisting-L2 access window on a stream's workspace: its lines are kept (hitProp persisting) while the streamed
// x / y columns of the same kernels pass through the rest of the cache -- the workspace is written by pass A,
// rewritten by pass B and read by pass C, and must not be evicted in between by 16 MiB of x and y per column pair.
static std::atomic<long long> g_l2_applied[64];     // persisting-L2 carve-out currently set on each device (bytes)
// The carve-out is a device-wide limit: while it is set, normal and streaming accesses of EVERY kernel lose that part
// of L2 (config 3 with 2 x 64 MiB workspaces: 9.3 k -> 6.4 k solves/s, config 5a 0.53 -> 0.69 ms, probe 5).  Products
// that do not use the window hand it back, together with the lines still tagged persisting.
static void release_l2_carveout() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_l2_applied[dev & 63].exchange(0) != 0) {
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
}
static void apply_l2_window(cudaStream_t st, void* base, size_t bytes, int nwindows) {
    const int mb = g_l2_persist_mb.load();
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    if (mb > 0 && base) {
        int dev = 0, max_persist = 0, max_window = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        size_t want = std::min<size_t>((size_t)mb << 20, (size_t)std::max(max_persist, 0));
        if (g_l2_applied[dev & 63].exchange((long long)want) != (long long)want) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        const size_t wbytes = std::min<size_t>(bytes, (size_t)std::max(max_window, 0));
        const double share = (double)want / (double)std::max<size_t>(wbytes * (size_t)std::max(nwindows, 1), 1);
        v.accessPolicyWindow.base_ptr = base;
        v.accessPolicyWindow.num_bytes = wbytes;
        v.accessPolicyWindow.hitRatio = (float)std::min(1.0, share);
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        v.accessPolicyWindow.num_bytes = 0;
        v.accessPolicyWindow.hitRatio = 0.f;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
}

static int ensure_buf(std::shared_ptr<DevBuf>& b, size_t bytes, const char* what, cudaStream_t s) {
    if (b && b->bytes >= bytes) return TB200_OK;
    if (b) { TB_TRY(retire_after(s)); b.reset(); }      // the old block may still be in use by work queued on s
    return dev_alloc_on(b, bytes, what, s);
}
// a product on a stream other than the one that produced the spectrum waits for it
// Once the event has been seen complete nothing is enqueued any more (and a stream under capture, which must not wait
// on uncaptured work of another stream, can use the handle: warm up -- or synchronize -- before capturing).
static int wait_ready(cudaEvent_t ev, std::atomic<bool>& done, cudaStream_t s, const char* what) {
    if (done.load(std::memory_order_acquire)) return TB200_OK;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
    if (cs != cudaStreamCaptureStatusNone)
        return set_error(TB200_CUDA_ERR, "%s: the factorization was produced on another stream and has not been seen complete; "
                         "run one product (or synchronize) before capturing this stream", what);
    const cudaError_t q = cudaEventQuery(ev);
    if (q == cudaSuccess) { done.store(true, std::memory_order_release); return TB200_OK; }
    if (q != cudaErrorNotReady) return cuda_status(q, what);
    TB_CUDA(cudaStreamWaitEvent(s, ev, 0), what);
    return TB200_OK;
}
static int wait_spectrum(tb200_fact* h, const void* spec, cudaStream_t s) {
    if (h->ev_spec && s != h->spec_stream) TB_TRY(wait_ready(h->ev_spec, h->spec_done, s, "spectrum ready"));
    if (spec && h->inv_spec && spec == h->inv_spec->p && h->ev_inv && s != h->inv_stream)
        TB_TRY(wait_ready(h->ev_inv, h->inv_done, s, "reciprocal spectrum ready"));
    return TB200_OK;
}
static int mark_spectrum_ready(tb200_fact* h, cudaStream_t s) {
    if (!h->ev_spec) TB_CUDA(cudaEventCreateWithFlags(&h->ev_spec, cudaEventDisableTiming), "event");
    h->spec_done.store(false);
    TB_CUDA(cudaEventRecord(h->ev_spec, s), "spectrum ready");
    h->spec_stream = s;
    h->spec_valid = true;
    return TB200_OK;
}

// One description of "apply a spectrum to columns": covers mul!, ldiv!, factorize's FFT and to_vc.
struct ConvCall {
    const void* x = nullptr; long long ldx = 0; int x_mode = IO_CPLX; long long n_in = 0; int reverse_x = 0;
    void* y = nullptr; long long ldy = 0; int y_mode = IO_CPLX; long long m_out = 0;
    long long ncols = 0, nfft = 0;
    const void* spec = nullptr;      // layout-order spectrum or null
    int do_fwd = 1, do_inv = 1;
    double scale = 1.0; double alpha[2] = {1, 0}, beta[2] = {0, 0}; int has_beta = 0;
    const int* colmap = nullptr;                    // device: user column of logical column j (NULL = identity)
    const double* col_alpha = nullptr; int col_alpha_stride = 0;   // device: per-user-column factor on alpha
    // the slots are the rows of an outer (three-level) transform: every slot has its own spectrum segment, and
    // forward-only / inverse-only calls run all of them (layout-order data is addressed slot by slot)
    int nested = 0;
};

// 3-D tensor map over a column-major matrix seen as [column][row of N2 elements][N2 elements], in 8-byte units
// (tma_cols.cuh).  Returns false when the driver entry point is missing or the layout is not expressible.
static bool make_cols_tensor_map(CUtensorMap* map, const void* base, int units_per_elem, long long N2, long long rows,
                                 long long ncols, long long ld_elems, int ta, int box_rows, int box_cols) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeFn>(p);
    }();
    if (!fn || rows <= 0 || ncols <= 0) return false;
    const size_t col_stride = (size_t)ld_elems * (size_t)units_per_elem * 8;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (col_stride & 15)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)(N2 * units_per_elem), (cuuint64_t)rows, (cuuint64_t)ncols};
    const cuuint64_t strides[2] = {(cuuint64_t)(N2 * units_per_elem * 8), (cuuint64_t)col_stride};
    const cuuint32_t box[3] = {(cuuint32_t)(ta * units_per_elem), (cuuint32_t)box_rows, (cuuint32_t)box_cols};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename R> static int run_conv3_t(tb200_fact* h, const ConvCall& c, cudaStream_t s);

template <typename R>
static int run_conv_t(tb200_fact* h, const ConvCall& c, cudaStream_t s) {
    if (h->three_level) return run_conv3_t<R>(h, c, s);
    ConvArgs<R> a;
    memset(&a, 0, sizeof(a));
    const cx<R>* tw_rows = nullptr;
    const cx<R>* tw_cols = nullptr;
    TB_TRY(stage_twiddles<R>(h->log2l2, &tw_rows));
    if (h->two_level) TB_TRY(stage_twiddles<R>(h->log2l1, &tw_cols));
    a.tw = tw_rows;
    a.x = c.x; a.ldx = c.ldx; a.x_mode = c.x_mode; a.n_in = c.n_in; a.reverse_x = c.reverse_x;
    a.y = c.y; a.ldy = c.ldy; a.y_mode = c.y_mode; a.m_out = c.m_out;
    a.ncols = c.ncols; a.nfft = c.nfft;
    a.spec = reinterpret_cast<const cx<R>*>(c.spec); a.spec_rows = 1;
    a.do_fwd = c.do_fwd; a.do_inv = c.do_inv;
    a.scale = (R)c.scale; a.alpha = mk<R>((R)c.alpha[0], (R)c.alpha[1]); a.beta = mk<R>((R)c.beta[0], (R)c.beta[1]);
    a.has_beta = c.has_beta;
    a.colmap = c.colmap; a.col_alpha = c.col_alpha; a.col_alpha_stride = c.col_alpha_stride;
    if (c.nfft <= 0) return TB200_OK;
    TB_TRY(wait_spectrum(h, c.spec, s));
    if (!h->two_level) {
        count_launch();
        PassTimer pt(3, s);
        return cuda_status(launch_rows<R>(h->log2l2, a, num_sms(), s), "k_rows");
    }
    // ---- two-level: A (strided cols, fwd) -> B (rows: fwd * spec inv) -> C (strided cols, inv)
    const long long Np = h->Np, N1 = 1ll << h->log2l1, N2 = 1ll << h->log2l2;
    long long G = std::max(1ll, g_group_bytes.load() / (long long)(Np * sizeof(cx<R>)));
    G = std::min(G, c.nfft);
    StreamWork* w = h->work_for(s);
    TB_TRY(ensure_buf(w->scratch, (size_t)G * Np * sizeof(cx<R>), "workspace", s));
    const int NSTR = g_overlap_streams.load();
    const bool overlap = NSTR >= 2 && c.nfft > G && c.do_fwd && c.do_inv;
    const int nstr = overlap ? NSTR : 1;
    cx<R>* scratch_of[tb200_fact::MAXS];
    cudaStream_t stream_of[tb200_fact::MAXS];
    for (int i = 0; i < tb200_fact::MAXS; ++i) { scratch_of[i] = reinterpret_cast<cx<R>*>(w->scratch->p); stream_of[i] = s; }
    if (overlap) {
        for (int i = 1; i < nstr; ++i) {
            TB_TRY(ensure_buf(w->scratch_extra[i], (size_t)G * Np * sizeof(cx<R>), "extra workspace", s));
            scratch_of[i] = reinterpret_cast<cx<R>*>(w->scratch_extra[i]->p);
        }
        for (int i = 0; i < nstr; ++i) {
            if (!w->hs[i]) {
                TB_CUDA(cudaStreamCreateWithFlags(&w->hs[i], cudaStreamNonBlocking), "helper stream");
                TB_CUDA(cudaEventCreateWithFlags(&w->ev_join[i], cudaEventDisableTiming), "event");
            }
            stream_of[i] = w->hs[i];
        }
        if (!w->ev_fork) TB_CUDA(cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming), "event");
        TB_CUDA(cudaEventRecord(w->ev_fork, s), "fork");
        for (int i = 0; i < nstr; ++i) TB_CUDA(cudaStreamWaitEvent(w->hs[i], w->ev_fork, 0), "fork");
    }
    cudaStream_t user_stream = s;
    // the window pays only when the workspaces in flight fit the carve-out as a whole (config 2: 2 x 32 MiB); larger
    // ones (2 x 64 MiB of config 3, 256 MiB of config 5) stream through DRAM anyway and a partial window costs them
    // 15-30 % (profiles/experiments_r02.md)
    const size_t ws_in_flight = (size_t)nstr * (size_t)G * Np * sizeof(cx<R>);
    // ... and only for multi-group products (with or without the stream overlap: pass C of a group finds its workspace
    // in L2 either way): a single transform (factorize, a lone vector) would flip the device-wide limit back and forth
    // between calls for nothing
    const bool multi_group = c.nfft > G && c.do_fwd && c.do_inv;
    // ... and when the spectrum every pass B streams fits beside them (config 3 on one stream: 64 MiB + 64 MiB, 8.2 k -> 7.9 k
    // solves/s with the window)
    const bool l2win = g_l2_persist_mb.load() > 0 && multi_group && ws_in_flight <= ((size_t)g_l2_persist_mb.load() << 20) &&
                       ws_in_flight + (size_t)Np * sizeof(cx<R>) <= ((size_t)104 << 20);
    if (!l2win && (multi_group || ws_in_flight > ((size_t)g_l2_persist_mb.load() << 20)))
        release_l2_carveout();              // a carve-out left behind by a smaller product halves the L2 this one sees
    for (int i = 0; i < nstr; ++i) {
        if (l2win) { apply_l2_window(stream_of[i], scratch_of[i], (size_t)G * Np * sizeof(cx<R>), nstr); w->win_set[i] = true; }
        else if (w->win_set[i]) { apply_l2_window(stream_of[i], nullptr, 0, 1); w->win_set[i] = false; }
    }
    a.log2n2 = h->log2l2;
    a.tw_lo = reinterpret_cast<const cx<R>*>(h->tw_lo->p);
    a.tw_hi = reinterpret_cast<const cx<R>*>(h->tw_hi->p);
    a.tw_hb = h->tw_hb;
    a.tw_g = reinterpret_cast<const cx<R>*>(h->tw_g->p);
    const size_t in_esz = (c.x_mode == IO_REAL || c.x_mode == IO_PAIR) ? sizeof(R) : sizeof(cx<R>);
    const size_t out_esz = (c.y_mode == IO_REAL || c.y_mode == IO_PAIR) ? sizeof(R) : sizeof(cx<R>);
    const long long cols_per_fft = (c.x_mode == IO_PAIR || c.y_mode == IO_PAIR) ? 2 : 1;
    // ---- optional TMA staging of the strided tiles (tma_cols.cuh) ----
    alignas(64) CUtensorMap xmap, ymap;
    bool tmaA = false, tmaC = false;
    TmaLaunch tlA{0, 0, false, false}, tlC{0, 0, false, false};
    if (g_tma.load() && !c.colmap && !c.col_alpha && sizeof(R) * 2 >= 8) {
        const long long halfN = Np / 2;
        auto tile_rows = [&](bool half) { return (int)std::min<long long>(half ? N1 / 2 : N1, 256); };
        if ((g_tma.load() & 1) && c.do_fwd && (c.x_mode == IO_CPLX || (c.x_mode == IO_PAIR && sizeof(R) == 8)) && c.n_in % N2 == 0) {
            const bool pair = c.x_mode == IO_PAIR;
            const bool half = g_prune_half && c.n_in <= halfN && h->log2l1 >= 4;
            if (cols_tma_supported<R>(h->log2l1, h->ta, pair) && (!c.reverse_x || c.n_in / N2 <= (half ? N1 / 2 : N1)) &&
                make_cols_tensor_map(&xmap, c.x, pair ? 1 : (int)(sizeof(cx<R>) / 8), N2, c.n_in / N2, c.ncols, c.ldx, h->ta,
                                     tile_rows(half), pair ? 2 : 1)) {
                tmaA = true;
                tlA = TmaLaunch{(int)(c.n_in / N2), 0, half, pair};
            }
        }
        if ((g_tma.load() & 2) && c.do_inv && !c.has_beta && (c.y_mode == IO_CPLX || (c.y_mode == IO_PAIR && sizeof(R) == 8)) && c.m_out % N2 == 0) {
            const bool pair = c.y_mode == IO_PAIR;
            const bool half = g_prune_half && c.m_out <= halfN && h->log2l1 >= 4;
            if (cols_tma_supported<R>(h->log2l1, h->ta, pair) &&
                make_cols_tensor_map(&ymap, c.y, pair ? 1 : (int)(sizeof(cx<R>) / 8), N2, c.m_out / N2, c.ncols, c.ldy, h->ta,
                                     tile_rows(half), pair ? 2 : 1)) {
                tmaC = true;
                tlC = TmaLaunch{(int)(c.m_out / N2), 0, half, pair};
            }
        }
    }
    // arguments of the strided passes of the group that starts at FFT slot f0 and holds g slots
    auto args_A = [&](long long f0, long long g, cx<R>* scratch) {
        ConvArgs<R> pa = a;
        pa.scratch = scratch;
        pa.nfft = g;
        pa.tw = tw_cols;
        pa.ncols = c.ncols - f0 * cols_per_fft;
        if (c.colmap) pa.colmap = c.colmap + f0 * cols_per_fft;
        else pa.x = reinterpret_cast<const char*>(c.x) + (size_t)f0 * cols_per_fft * c.ldx * in_esz;
        return pa;
    };
    auto args_C = [&](long long f0, long long g, cx<R>* scratch) {
        ConvArgs<R> pc = a;
        pc.scratch = scratch;
        pc.nfft = g;
        pc.tw = tw_cols;
        pc.ncols = c.ncols - f0 * cols_per_fft;
        if (c.colmap) pc.colmap = c.colmap + f0 * cols_per_fft;
        else {
            pc.y = reinterpret_cast<char*>(c.y) + (size_t)f0 * cols_per_fft * c.ldy * out_esz;
            if (c.col_alpha) pc.col_alpha = c.col_alpha + f0 * cols_per_fft * c.col_alpha_stride;   // indexed by user column
        }
        return pc;
    };
    // Fused C+A (k_colsCA): the last pass of a group and the first pass of the next group on the same stream / workspace
    // run as one launch.  Both sides must agree on the half-padding variant.
    const long long halfN = Np / 2;
    const bool half_ok = g_prune_half && le_of<R>() == 4 && h->log2l1 >= 4;
    const bool halfA = half_ok && c.n_in <= halfN && c.x_mode != IO_SPEC;
    const bool halfC = half_ok && c.m_out <= halfN;
#ifdef TB_EXP_CA
    const bool fuse = g_ca_fuse.load() && c.do_fwd && c.do_inv && !tmaA && !tmaC && halfA == halfC &&
                      c.nfft > (long long)nstr * G && cols_ca_supported<R>(h->log2l1, h->ta);
    const bool stage = fuse && g_ca_fuse.load() >= 2 && halfA && (c.x_mode == IO_PAIR || c.x_mode == IO_REAL);
#else
    const bool fuse = false;        // experiment build only (csrc/experiments/ca_kernels.cuh): measured slower
    (void)halfA; (void)halfC;
#endif
    // PDL hides the launch latency between the passes of one product; the per-pass timers bracket launches with events,
    // which would serialise them again, so the two are exclusive
    const bool use_pdl = !g_time_passes.load() && !tmaA && !tmaC && (g_pdl.load() >= 2 || (g_pdl.load() == 1 && !overlap));
    long long group_index = 0;
    for (long long f0 = 0; f0 < c.nfft; f0 += G, ++group_index) {
        const long long g = std::min(G, c.nfft - f0);
        cx<R>* scratch = scratch_of[group_index % nstr];
        s = stream_of[group_index % nstr];
        a.scratch = scratch;
        // pass A
        if (c.do_fwd) {
            if (c.x_mode == IO_SPEC) return set_error(TB200_BAD_ARG, "internal: IO_SPEC input with forward transform");
            const ConvArgs<R> pa = args_A(f0, g, scratch);
            count_launch();
            if (fuse && group_index >= nstr) {       // with pass C of the previous group of this stream (always a full group)
                const ConvArgs<R> pc = args_C(f0 - (long long)nstr * G, G, scratch);
                PassTimer pt(4, s);
#ifdef TB_EXP_CA
                TB_CUDA(launch_colsCA<R>(h->log2l1, h->ta, pc, pa, halfA, stage, s), "k_colsCA");
#else
                (void)pc;
#endif
            } else {
                PassTimer pt(0, s);
                if (tmaA) {
                    tlA.col0 = (int)(f0 * cols_per_fft);
                    TB_CUDA(launch_colsA_tma<R>(h->log2l1, h->ta, pa, &xmap, tlA, s), "k_colsA_tma");
                } else
                TB_CUDA(launch_colsA<R>(h->log2l1, h->ta, pa, s), "k_colsA");
            }
        }
        // pass B: rows of the scratch, in place
        {
            ConvArgs<R> pb = a;
            pb.nfft = g * N1; pb.ncols = g * N1;
            pb.n_in = N2; pb.reverse_x = 0; pb.m_out = N2;
            pb.ldx = N2; pb.ldy = N2;
            pb.x = scratch; pb.x_mode = IO_CPLX; pb.y = scratch; pb.y_mode = IO_CPLX;
            pb.spec_rows = (int)N1;
            if (c.nested && pb.spec) {   // rows of an outer transform: slot f owns the spectrum segment f
                pb.spec = pb.spec + (size_t)f0 * Np;
                pb.spec_rows = (int)(g * N1);
            }
            pb.scale = (R)1; pb.alpha = mk<R>(1, 0); pb.beta = mk<R>(0, 0); pb.has_beta = 0;
            pb.colmap = nullptr; pb.col_alpha = nullptr;
            if (g_rows_direct.load()) { pb.x_mode = IO_SPEC; pb.y_mode = IO_SPEC; }   // full rows, unit epilogue: the plain
                                                        // load / store path of k_rows (same addresses, no per-element bounds and mode logic)
            if (!c.do_fwd) {             // inverse only: rows come from a layout-order spectrum
                pb.x = c.nested ? (const void*)(reinterpret_cast<const cx<R>*>(c.x) + (size_t)f0 * Np) : c.x;
                pb.x_mode = IO_SPEC; pb.spec = nullptr;
                if (!c.nested && (g != 1 || c.nfft != 1)) return set_error(TB200_BAD_ARG, "internal: inverse-only runs one transform");
            }
            if (!c.do_inv) {             // forward only: rows go out in layout order
                pb.y = c.nested ? (void*)(reinterpret_cast<cx<R>*>(c.y) + (size_t)f0 * Np) : c.y;
                pb.y_mode = IO_SPEC; pb.spec = nullptr;
                if (!c.nested && (g != 1 || c.nfft != 1)) return set_error(TB200_BAD_ARG, "internal: forward-only runs one transform");
            }
            pb.pdl = (use_pdl && c.do_fwd) ? 1 : 0;          // behind pass A of this group
            count_launch();
            PassTimer pt(1, s);
            TB_CUDA(launch_rows<R>(h->log2l2, pb, num_sms(), s), "k_rows (pass B)");
        }
        // pass C (deferred into the next group's fused launch when this stream has one)
        if (c.do_inv && !(fuse && f0 + (long long)nstr * G < c.nfft)) {
            ConvArgs<R> pc = args_C(f0, g, scratch);
            pc.pdl = use_pdl ? 1 : 0;                          // behind pass B of this group
            count_launch();
            PassTimer pt(2, s);
            if (tmaC) {
                tlC.col0 = (int)(f0 * cols_per_fft);
                TB_CUDA(launch_colsC_tma<R>(h->log2l1, h->ta, pc, &ymap, tlC, s), "k_colsC_tma");
            } else
            TB_CUDA(launch_colsC<R>(h->log2l1, h->ta, pc, s), "k_colsC");
        }
    }
    if (overlap) {
        for (int i = 0; i < nstr; ++i) {
            TB_CUDA(cudaEventRecord(w->ev_join[i], w->hs[i]), "join");
            TB_CUDA(cudaStreamWaitEvent(user_stream, w->ev_join[i], 0), "join");
        }
    }
    return TB200_OK;
}

// Three-level transform: N' = N0 x Nr.  Per FFT slot: the outer strided pass (k_colsA, N0-point columns at stride Nr)
// fills a workspace of N' points; its N0 rows are then ONE batched call of the two-level machinery on the child plan
// (forward -> x spectrum segment of the row -> inverse, in place, column groups / overlap streams / L2 window as for
// any batch); the outer inverse pass (k_colsC) applies the epilogue.  Forward-only (factorize) and inverse-only
// (spectrum -> coefficients) calls run the matching halves.
template <typename R>
static int run_conv3_t(tb200_fact* h, const ConvCall& c, cudaStream_t s) {
    if (c.nfft <= 0) return TB200_OK;
    if (c.colmap || c.col_alpha) return set_error(TB200_UNSUPPORTED, "column selection is not available on the three-level plan");
    tb200_fact* hi = h->inner.get();
    const long long Np = h->Np, N0 = 1ll << h->log2l1, Nr = 1ll << h->log2l2;
    const cx<R>* tw_cols = nullptr;
    TB_TRY(stage_twiddles<R>(h->log2l1, &tw_cols));
    TB_TRY(wait_spectrum(h, c.spec, s));
    StreamWork* w = h->work_for(s);
    TB_TRY(ensure_buf(w->scratch, (size_t)Np * sizeof(cx<R>), "outer workspace", s));
    cx<R>* W = reinterpret_cast<cx<R>*>(w->scratch->p);
    ConvArgs<R> a;
    memset(&a, 0, sizeof(a));
    a.x = c.x; a.ldx = c.ldx; a.x_mode = c.x_mode; a.n_in = c.n_in; a.reverse_x = c.reverse_x;
    a.y = c.y; a.ldy = c.ldy; a.y_mode = c.y_mode; a.m_out = c.m_out;
    a.ncols = c.ncols; a.nfft = 1;
    a.do_fwd = c.do_fwd; a.do_inv = c.do_inv;
    a.scale = (R)c.scale; a.alpha = mk<R>((R)c.alpha[0], (R)c.alpha[1]); a.beta = mk<R>((R)c.beta[0], (R)c.beta[1]);
    a.has_beta = c.has_beta;
    a.tw = tw_cols;
    a.scratch = W;
    a.log2n2 = h->log2l2;
    a.tw_lo = reinterpret_cast<const cx<R>*>(h->tw_lo->p);
    a.tw_hi = reinterpret_cast<const cx<R>*>(h->tw_hi->p);
    a.tw_hb = h->tw_hb;
    a.tw_g = nullptr;
    const size_t in_esz = (c.x_mode == IO_REAL || c.x_mode == IO_PAIR) ? sizeof(R) : sizeof(cx<R>);
    const size_t out_esz = (c.y_mode == IO_REAL || c.y_mode == IO_PAIR) ? sizeof(R) : sizeof(cx<R>);
    const long long cols_per_fft = (c.x_mode == IO_PAIR || c.y_mode == IO_PAIR) ? 2 : 1;
    if ((!c.do_fwd || !c.do_inv) && c.nfft != 1) return set_error(TB200_BAD_ARG, "internal: one-directional runs take one transform");
    for (long long f = 0; f < c.nfft; ++f) {
        if (c.do_fwd) {
            ConvArgs<R> pa = a;
            pa.ncols = c.ncols - f * cols_per_fft;
            pa.x = reinterpret_cast<const char*>(c.x) + (size_t)f * cols_per_fft * c.ldx * in_esz;
            count_launch();
            PassTimer pt(0, s);
            TB_CUDA(launch_colsA<R>(h->log2l1, h->ta, pa, s), "k_colsA (outer)");
        }
        ConvCall ci;
        ci.x = c.do_fwd ? (const void*)W : c.x; ci.ldx = Nr; ci.x_mode = IO_CPLX; ci.n_in = Nr;
        ci.y = c.do_inv ? (void*)W : c.y; ci.ldy = Nr; ci.y_mode = IO_CPLX; ci.m_out = Nr;
        ci.ncols = N0; ci.nfft = N0;
        ci.spec = (c.do_fwd && c.do_inv) ? c.spec : nullptr;
        ci.do_fwd = c.do_fwd; ci.do_inv = c.do_inv;
        ci.nested = 1;
        TB_TRY(run_conv_t<R>(hi, ci, s));
        if (c.do_inv) {
            ConvArgs<R> pc = a;
            pc.ncols = c.ncols - f * cols_per_fft;
            pc.y = reinterpret_cast<char*>(c.y) + (size_t)f * cols_per_fft * c.ldy * out_esz;
            count_launch();
            PassTimer pt(2, s);
            TB_CUDA(launch_colsC<R>(h->log2l1, h->ta, pc, s), "k_colsC (outer)");
        }
    }
    return TB200_OK;
}

static int run_conv(tb200_fact* h, const ConvCall& c, cudaStream_t s) {
    return h->prec ? run_conv_t<double>(h, c, s) : run_conv_t<float>(h, c, s);
}

// ---------------------------------------------------------------------------------------
// factorize
// ---------------------------------------------------------------------------------------
static int check_factorize_args(int kind, int dtype, int64_t m, int64_t n, const void* vc, const void* vr, bool need_ptrs) {
    if (kind < 0 || kind > 5) return set_error(TB200_BAD_ARG, "unknown matrix kind %d", kind);
    if (!dt_valid(dtype)) return set_error(TB200_BAD_ARG, "unknown dtype %d", dtype);
    if (m <= 0 || n <= 0) return set_error(TB200_BAD_ARG, "matrix size must be positive, got %lldx%lld", (long long)m, (long long)n);
    if (kind != TB200_TOEPLITZ && kind != TB200_HANKEL && m != n)
        return set_error(TB200_DIM_MISMATCH, "matrix kind %d must be square, got %lldx%lld", kind, (long long)m, (long long)n);
    if (need_ptrs) {
        if (!vc) return set_error(TB200_BAD_ARG, "vc is NULL");
        if (kind == TB200_TOEPLITZ && !vr) return set_error(TB200_BAD_ARG, "vr is NULL for a general Toeplitz matrix");
    }
    return TB200_OK;
}

static int new_handle(int kind, int dtype, int64_t m, int64_t n, cudaStream_t s, tb200_fact** out) {
    std::unique_ptr<tb200_fact> h(new tb200_fact());
    h->kind = kind; h->dtype = dtype; h->prec = dt_prec(dtype); h->a_real = !dt_is_cplx(dtype);
    h->m = m; h->n = n;
    TB_CUDA(cudaGetDevice(&h->device), "cudaGetDevice");
    if (kind == TB200_CIRCULANT) {
        h->Np = n;
        if ((n & (n - 1)) != 0 || n < 16) {       // exact n-point DFT semantics through a chirp-z plan
            h->bluestein = true;
            h->reverse_x = 0;
            TB_TRY(dev_alloc_on(h->spec, (size_t)n * (h->prec ? 16 : 8), "spectrum", s));
            *out = h.release();
            return TB200_OK;
        }
    } else {
        h->Np = 1ll << next_log2(m + n - 1);
    }
    if (h->Np < 16) h->Np = 16;
    h->log2np = next_log2(h->Np);
    h->reverse_x = (kind == TB200_HANKEL);
    TB_TRY(make_plan(h.get()));
    TB_TRY(dev_alloc_on(h->spec, (size_t)h->Np * (h->prec ? 16 : 8), "spectrum", s));
    *out = h.release();
    return TB200_OK;
}

template <typename R>
static int factorize_t(tb200_fact* h, const void* vc, const void* vr, cudaStream_t s) {
    // K4: embedding into the workspace, then forward FFT into the layout-order spectrum
    StreamWork* w = h->work_for(s);
    TB_TRY(ensure_buf(w->scratch, (size_t)h->Np * sizeof(cx<R>), "workspace", s));
    cx<R>* c = reinterpret_cast<cx<R>*>(w->scratch->p);
    k_embed<R><<<ew_grid(h->Np), 256, 0, s>>>(c, h->Np, h->kind, vc, vr, !h->a_real, h->m, h->n);
    count_launch();
    TB_CUDA(cudaGetLastError(), "k_embed");
    ConvCall cc;
    cc.x = c; cc.ldx = h->Np; cc.x_mode = IO_CPLX; cc.n_in = h->Np;
    cc.y = h->spec->p; cc.ldy = h->Np; cc.y_mode = IO_SPEC; cc.m_out = h->Np;
    cc.ncols = 1; cc.nfft = 1; cc.do_fwd = 1; cc.do_inv = 0;
    TB_TRY(run_conv_t<R>(h, cc, s));
    return mark_spectrum_ready(h, s);
}

static int mode_for(bool cplx) { return cplx ? IO_CPLX : IO_REAL; }

// ---------------------------------------------------------------------------------------
// Bluestein plan: exact n-point DFTs for Circulants of any size on top of the pow2 kernels
// ---------------------------------------------------------------------------------------
template <typename R>
static int bs_setup_t(tb200_fact* h, cudaStream_t s) {
    const int64_t n = h->n;
    std::vector<cx<R>> c((size_t)n), cc((size_t)n);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int64_t j = 0; j < n; ++j) {
        const long long q = (long long)((j * j) % (2 * n));          // exact: j^2 < 2^50
        const long double ang = pi * (long double)q / (long double)n;
        c[(size_t)j] = mk<R>((R)cosl(ang), (R)(-sinl(ang)));
        cc[(size_t)j] = mk<R>((R)cosl(ang), (R)sinl(ang));
    }
    TB_TRY(dev_alloc_on(h->bs_chirp, sizeof(cx<R>) * (size_t)n, "chirp", s));
    TB_CUDA(cudaMemcpyAsync(h->bs_chirp->p, c.data(), sizeof(cx<R>) * (size_t)n, cudaMemcpyHostToDevice, s), "chirp upload");
    std::shared_ptr<DevBuf> tvc;
    TB_TRY(dev_alloc_on(tvc, sizeof(cx<R>) * (size_t)n, "chirp filter", s));
    TB_CUDA(cudaMemcpyAsync(tvc->p, cc.data(), sizeof(cx<R>) * (size_t)n, cudaMemcpyHostToDevice, s), "chirp filter upload");
    tb200_fact* t = nullptr;
    TB_TRY(new_handle(TB200_SYMMETRIC, h->prec ? TB200_C64 : TB200_C32, n, n, s, &t));
    h->bs_conv.reset(t);
    TB_TRY(factorize_t<R>(t, tvc->p, nullptr, s));
    TB_CUDA(cudaStreamSynchronize(s), "bluestein setup sync");           // host staging vectors go out of scope
    return TB200_OK;
}

template <typename R>
static int bs_conv_t(tb200_fact* h, cx<R>* v, const cx<R>* u, int64_t ld, int64_t ncols, cudaStream_t s) {
    tb200_fact* t = h->bs_conv.get();
    ConvCall c;
    c.x = u; c.ldx = ld; c.x_mode = IO_CPLX; c.n_in = h->n;
    c.y = v; c.ldy = ld; c.y_mode = IO_CPLX; c.m_out = h->n;
    c.ncols = ncols; c.nfft = ncols; c.spec = t->spec->p;
    c.scale = 1.0 / (double)t->Np;
    return run_conv_t<R>(t, c, s);
}

// y = alpha * maybereal(OP(x)) + beta*y with OP = forward DFT (mode 0), inverse DFT (mode 1, x is a
// natural-order spectrum) or circulant apply ifft(fft(x) .* spec) (mode 2), column by column chunk.
template <typename R>
static int bs_run_t(tb200_fact* h, int mode, const void* spec, void* y, int64_t ldy, int y_cplx, int take_real,
                    const void* x, int64_t ldx, int x_cplx, int64_t ncols, const double alpha[2], const double beta[2],
                    cudaStream_t s) {
    const int64_t n = h->n;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(ncols, (64ll << 20) / (int64_t)(n * sizeof(cx<R>))));
    const size_t wbytes = (size_t)chunk * n * sizeof(cx<R>);
    StreamWork* w = h->work_for(s);
    TB_TRY(ensure_buf(w->bs_u, wbytes, "bluestein workspace", s));
    TB_TRY(ensure_buf(w->bs_v, wbytes, "bluestein workspace", s));
    cx<R>* u = (cx<R>*)w->bs_u->p; cx<R>* v = (cx<R>*)w->bs_v->p;
    const cx<R>* chirp = (const cx<R>*)h->bs_chirp->p;
    const cx<R> al = mk<R>((R)(alpha ? alpha[0] : 1.0), (R)(alpha ? alpha[1] : 0.0));
    const cx<R> be = mk<R>((R)(beta ? beta[0] : 0.0), (R)(beta ? beta[1] : 0.0));
    const int has_beta = (be.x != 0 || be.y != 0);
    const size_t xsz = x_cplx ? sizeof(cx<R>) : sizeof(R), ysz = y_cplx ? sizeof(cx<R>) : sizeof(R);
    for (int64_t c0 = 0; c0 < ncols; c0 += chunk) {
        const int64_t nc = std::min(chunk, ncols - c0);
        const void* xs = (const char*)x + (size_t)c0 * ldx * xsz;
        void* ys = (char*)y + (size_t)c0 * ldy * ysz;
        const unsigned grid = ew_grid(n * nc);
        k_bs_pre<R><<<grid, 256, 0, s>>>(u, n, xs, ldx, x_cplx, mode == 1, chirp, n, nc);
        count_launch();
        TB_TRY(bs_conv_t<R>(h, v, u, n, nc, s));
        if (mode == 2) {
            k_bs_mid<R><<<grid, 256, 0, s>>>(u, v, n, chirp, (const cx<R>*)spec, n, nc);
            count_launch();
            TB_TRY(bs_conv_t<R>(h, v, u, n, nc, s));
        }
        const int inv = (mode != 0);
        k_bs_post<R><<<grid, 256, 0, s>>>(ys, ldy, y_cplx, v, n, chirp, n, nc, inv ? (R)(1.0 / (double)n) : (R)1, inv,
                                          take_real, al, be, has_beta);
        count_launch();
        TB_CUDA(cudaGetLastError(), "bluestein kernels");
    }
    return TB200_OK;
}

static int bs_run(tb200_fact* h, int mode, const void* spec, void* y, int64_t ldy, int y_cplx, int take_real, const void* x,
                  int64_t ldx, int x_cplx, int64_t ncols, const double alpha[2], const double beta[2], cudaStream_t s) {
    return h->prec ? bs_run_t<double>(h, mode, spec, y, ldy, y_cplx, take_real, x, ldx, x_cplx, ncols, alpha, beta, s)
                   : bs_run_t<float>(h, mode, spec, y, ldy, y_cplx, take_real, x, ldx, x_cplx, ncols, alpha, beta, s);
}

// factorize for a Bluestein Circulant: spectrum = DFT_n(vc), natural order
static int bs_factorize(tb200_fact* h, const void* vc, cudaStream_t s) {
    TB_TRY(h->prec ? bs_setup_t<double>(h, s) : bs_setup_t<float>(h, s));
    TB_TRY(bs_run(h, 0, nullptr, h->spec->p, h->n, 1, 0, vc, h->n, !h->a_real, 1, nullptr, nullptr, s));
    return mark_spectrum_ready(h, s);
}

}  // namespace tb

extern "C" {

const char* tb200_version(void) { return "toeplitz_b200 0.1.0 (sm_100a)"; }
const char* tb200_last_error(void) { return g_err; }
int64_t tb200_launch_count(void) { return g_launches.load(); }

int tb200_pass_times_ex(double* ms, int64_t* counts, int n) {
    if (!ms || !counts || n < 1) return set_error(TB200_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lock(g_pass_mutex);
    for (int w = 0; w < 5; ++w) {
        double tot = 0; int64_t cnt = 0;
        for (auto& pr : g_pass[w].ev) {
            float t = 0;
            cudaEventSynchronize(pr.second);
            if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) { tot += t; cnt += 1; }
            cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
        }
        g_pass[w].ev.clear();
        if (w < n) { ms[w] = tot; counts[w] = cnt; }
        else { ms[n - 1] += tot; counts[n - 1] += cnt; }       // a shorter array folds the rest into its last entry
    }
    return TB200_OK;
}

int tb200_pass_times(double ms[4], int64_t counts[4]) {
    // four-entry form: the fused C+A launches (entry 4 of tb200_pass_times_ex) are reported with pass C
    double m5[5]; int64_t c5[5];
    TB_TRY(tb200_pass_times_ex(m5, c5, 5));
    if (!ms || !counts) return set_error(TB200_BAD_ARG, "NULL argument");
    for (int w = 0; w < 4; ++w) { ms[w] = m5[w]; counts[w] = c5[w]; }
    ms[2] += m5[4]; counts[2] += c5[4];
    return TB200_OK;
}

int tb200_set_option(const char* name, int64_t value) {
    if (!name) return set_error(TB200_BAD_ARG, "option name is NULL");
    if (!strcmp(name, "group_bytes")) { g_group_bytes = std::max<int64_t>(value, 1 << 20); return TB200_OK; }
    if (!strcmp(name, "single_max_f64")) { g_single_max_f64 = (int)std::min<int64_t>(std::max<int64_t>(value, 4), 13); return TB200_OK; }
    if (!strcmp(name, "single_max_f32")) { g_single_max_f32 = (int)std::min<int64_t>(std::max<int64_t>(value, 4), 14); return TB200_OK; }
    if (!strcmp(name, "krylov_work_bytes")) { g_krylov_work_bytes = std::max<int64_t>(value, 1 << 20); return TB200_OK; }
    if (!strcmp(name, "host_chunk_bytes")) { g_host_chunk_bytes = std::max<int64_t>(value, 1 << 20); return TB200_OK; }
    if (!strcmp(name, "time_passes")) { g_time_passes = value ? 1 : 0; return TB200_OK; }
    if (!strcmp(name, "ca_fuse")) { g_ca_fuse = (int)std::min<int64_t>(std::max<int64_t>(value, 0), 2); return TB200_OK; }
    if (!strcmp(name, "pdl")) { g_pdl = (int)std::min<int64_t>(std::max<int64_t>(value, 0), 2); return TB200_OK; }
    if (!strcmp(name, "l2_persist_mb")) { g_l2_persist_mb = (int)std::min<int64_t>(std::max<int64_t>(value, 0), 1024); return TB200_OK; }
    if (!strcmp(name, "lev_gen")) { g_lev_gen = (int)(value & 1); return TB200_OK; }
    if (!strcmp(name, "prune_half")) { g_prune_half = value ? 1 : 0; return TB200_OK; }
    if (!strcmp(name, "rows_direct")) { g_rows_direct = value != 0; return TB200_OK; }
    if (!strcmp(name, "pair_columns")) { g_pair_columns = value != 0; return TB200_OK; }
    if (!strcmp(name, "tma")) { g_tma = (int)std::min<int64_t>(std::max<int64_t>(value, 0), 3); return TB200_OK; }   // bit 0: pass A loads, bit 1: pass C stores
    if (!strcmp(name, "overlap_streams")) { g_overlap_streams = (int)std::min<int64_t>(std::max<int64_t>(value, 0), tb200_fact::MAXS); return TB200_OK; }
    if (!strcmp(name, "l2max_f64")) { g_l2max_f64 = (int)value; return TB200_OK; }
    if (!strcmp(name, "l2max_f32")) { g_l2max_f32 = (int)value; return TB200_OK; }
    if (!strcmp(name, "ta_cap")) { g_ta_cap = (int)value; return TB200_OK; }
    if (!strcmp(name, "force_l1")) { g_force_l1 = (int)value; return TB200_OK; }
    return set_error(TB200_BAD_ARG, "unknown option '%s'", name);
}

int tb200_factorize_empty(int kind, int dtype, int64_t m, int64_t n, void* stream, tb200_handle_t* out) {
    (void)stream;
    if (!out) return set_error(TB200_BAD_ARG, "out is NULL");
    TB_TRY(check_factorize_args(kind, dtype, m, n, nullptr, nullptr, false));
    tb200_fact* h = nullptr;
    TB_TRY(new_handle(kind, dtype, m, n, (cudaStream_t)stream, &h));
    if (h->bluestein) {
        int st = h->prec ? bs_setup_t<double>(h, (cudaStream_t)stream) : bs_setup_t<float>(h, (cudaStream_t)stream);
        if (st != TB200_OK) { delete h; return st; }
    }
    // the caller fills the spectrum buffer (tb200_bcast_spectrum, or its own collective followed by
    // tb200_spectrum_ready on the stream that wrote it)
    h->spec_valid = true;
    *out = h;
    return TB200_OK;
}

int tb200_factorize(int kind, int dtype, int64_t m, int64_t n, const void* vc, const void* vr, void* stream,
                    tb200_handle_t* out) {
    if (!out) return set_error(TB200_BAD_ARG, "out is NULL");
    TB_TRY(check_factorize_args(kind, dtype, m, n, vc, vr, true));
    tb200_fact* h = nullptr;
    TB_TRY(new_handle(kind, dtype, m, n, (cudaStream_t)stream, &h));
    int st = h->bluestein ? bs_factorize(h, vc, (cudaStream_t)stream)
           : h->prec ? factorize_t<double>(h, vc, vr, (cudaStream_t)stream) : factorize_t<float>(h, vc, vr, (cudaStream_t)stream);
    if (st != TB200_OK) { delete h; return st; }
    *out = h;
    return TB200_OK;
}

int tb200_factorize_host(int kind, int dtype, int64_t m, int64_t n, const void* vc_host, const void* vr_host,
                         void* stream, tb200_handle_t* out) {
    TB_TRY(check_factorize_args(kind, dtype, m, n, vc_host, vr_host, true));
    const size_t esz = dt_size(dtype);
    const int64_t nvc = (kind == TB200_HANKEL) ? m + n - 1 : (kind == TB200_TOEPLITZ ? m : n);
    const int64_t nvr = (kind == TB200_TOEPLITZ) ? n : 0;
    std::shared_ptr<DevBuf> dvc, dvr;
    cudaStream_t s = (cudaStream_t)stream;
    TB_TRY(dev_alloc_on(dvc, (size_t)nvc * esz, "vc staging", s));
    TB_CUDA(cudaMemcpyAsync(dvc->p, vc_host, (size_t)nvc * esz, cudaMemcpyHostToDevice, s), "vc upload");
    if (nvr) {
        TB_TRY(dev_alloc_on(dvr, (size_t)nvr * esz, "vr staging", s));
        TB_CUDA(cudaMemcpyAsync(dvr->p, vr_host, (size_t)nvr * esz, cudaMemcpyHostToDevice, s), "vr upload");
    }
    int st = tb200_factorize(kind, dtype, m, n, dvc->p, nvr ? dvr->p : nullptr, stream, out);
    TB_CUDA(cudaStreamSynchronize(s), "factorize_host sync");   // staging buffers are freed on return
    return st;
}

int tb200_destroy(tb200_handle_t h) {
    DeviceGuard dev_guard(h);
    delete h;
    return TB200_OK;
}

int tb200_info(tb200_handle_t h, int64_t* m, int64_t* n, int64_t* n_embed, int* dtype, int* kind) {
    if (!h) return set_error(TB200_BAD_ARG, "handle is NULL");
    if (m) *m = h->m;
    if (n) *n = h->n;
    if (n_embed) *n_embed = h->Np;
    if (dtype) *dtype = h->dtype;
    if (kind) *kind = h->kind;
    return TB200_OK;
}

int tb200_spectrum_ready(tb200_handle_t h, void* stream) {
    DeviceGuard dev_guard(h);
    if (!h) return set_error(TB200_BAD_ARG, "handle is NULL");
    return mark_spectrum_ready(h, (cudaStream_t)stream);
}

int tb200_spectrum_buffer(tb200_handle_t h, void** ptr, int64_t* bytes) {
    if (!h || !ptr || !bytes) return set_error(TB200_BAD_ARG, "NULL argument");
    *ptr = h->spec->p;
    *bytes = (int64_t)h->Np * (h->prec ? 16 : 8);
    return TB200_OK;
}

// ---- mul! --------------------------------------------------------------------------------
struct ColSel {            // optional column selection / per-column scalars of the batched Krylov drivers (device pointers)
    const int* colmap = nullptr; const double* col_alpha = nullptr; int col_alpha_stride = 0;
};
static int mul_impl(tb200_fact* h, const void* spec, bool is_ldiv, void* y, int64_t ldy, int ydtype, const void* x,
                    int64_t ldx, int xdtype, int64_t ncols, const double alpha[2], const double beta[2], cudaStream_t s,
                    const ColSel* sel = nullptr) {
    if (!h) return set_error(TB200_BAD_ARG, "handle is NULL");
    if (!dt_valid(xdtype) || !dt_valid(ydtype)) return set_error(TB200_BAD_ARG, "unknown dtype");
    if (dt_prec(xdtype) != h->prec || dt_prec(ydtype) != h->prec)
        return set_error(TB200_BAD_ARG, "vector precision must match the factorization's (%s)", h->prec ? "Float64" : "Float32");
    if (ncols < 0) return set_error(TB200_BAD_ARG, "negative column count");
    if (ncols == 0) return TB200_OK;
    if (!x || !y) return set_error(TB200_BAD_ARG, "x or y is NULL");
    if (ldx < h->n || ldy < h->m)
        return set_error(TB200_DIM_MISMATCH, "leading dimensions (%lld, %lld) smaller than the matrix size %lldx%lld",
                         (long long)ldy, (long long)ldx, (long long)h->m, (long long)h->n);
    const bool xc = dt_is_cplx(xdtype), yc = dt_is_cplx(ydtype);
    if (h->bluestein) {
        if (sel && (sel->colmap || sel->col_alpha)) return set_error(TB200_UNSUPPORTED, "internal: column selection on a Bluestein plan");
        if (!yc && (!h->a_real || (!is_ldiv && (xc || (alpha && alpha[1] != 0.0) || (beta && beta[1] != 0.0)))))
            return set_error(TB200_BAD_ARG, "a real y requires a real matrix, real x and real alpha/beta (InexactError in the reference)");
        return bs_run(h, 2, spec, y, ldy, yc, 0, x, ldx, xc, ncols, alpha, beta, s);
    }
    ConvCall c;
    c.x = x; c.ldx = ldx; c.n_in = h->n; c.reverse_x = h->reverse_x;
    c.y = y; c.ldy = ldy; c.m_out = h->m;
    c.ncols = ncols; c.spec = spec;
    c.scale = 1.0 / (double)h->Np;
    c.alpha[0] = alpha ? alpha[0] : 1.0; c.alpha[1] = alpha ? alpha[1] : 0.0;
    c.beta[0] = beta ? beta[0] : 0.0; c.beta[1] = beta ? beta[1] : 0.0;
    c.has_beta = (c.beta[0] != 0.0 || c.beta[1] != 0.0);
    if (sel) { c.colmap = sel->colmap; c.col_alpha = sel->col_alpha; c.col_alpha_stride = sel->col_alpha_stride; }
    if (!yc) {
        if (!is_ldiv && (!h->a_real || xc || c.alpha[1] != 0.0 || c.beta[1] != 0.0))
            return set_error(TB200_BAD_ARG, "a real y requires a real matrix, real x and real alpha/beta (InexactError in the reference)");
        if (is_ldiv && !h->a_real)
            return set_error(TB200_BAD_ARG, "real b with a complex Circulant (InexactError in the reference)");
        if (g_pair_columns.load()) { c.x_mode = IO_PAIR; c.y_mode = IO_PAIR; c.nfft = (ncols + 1) / 2; }
        else { c.x_mode = IO_REAL; c.y_mode = IO_REAL; c.nfft = ncols; }
    } else {
        c.x_mode = mode_for(xc); c.y_mode = IO_CPLX; c.nfft = ncols;
    }
    return run_conv(h, c, s);
}

int tb200_mul(tb200_handle_t h, void* y, int64_t ldy, int ydtype, const void* x, int64_t ldx, int xdtype, int64_t ncols,
              const double alpha[2], const double beta[2], void* stream) {
    DeviceGuard dev_guard(h);
    if (!h) return set_error(TB200_BAD_ARG, "handle is NULL");
    return mul_impl(h, h->spec->p, false, y, ldy, ydtype, x, ldx, xdtype, ncols, alpha, beta, (cudaStream_t)stream);
}

int tb200_mul_host(tb200_handle_t h, void* y_host, int64_t ldy, int ydtype, const void* x_host, int64_t ldx, int xdtype,
                   int64_t ncols, const double alpha[2], const double beta[2]) {
    DeviceGuard dev_guard(h);
    if (!h) return set_error(TB200_BAD_ARG, "handle is NULL");
    if (!dt_valid(xdtype) || !dt_valid(ydtype)) return set_error(TB200_BAD_ARG, "unknown dtype");
    if (ncols <= 0) return TB200_OK;
    if (ldx < h->n || ldy < h->m) return set_error(TB200_DIM_MISMATCH, "leading dimensions smaller than the matrix size");
    const size_t xsz = dt_size(xdtype), ysz = dt_size(ydtype);
    const bool has_beta = beta && (beta[0] != 0.0 || beta[1] != 0.0);
    // column chunks (even, so real columns keep pairing up)
    int64_t cc = std::max<int64_t>(2, g_host_chunk_bytes.load() / (int64_t)std::max<size_t>(h->n * xsz, h->m * ysz));
    cc &= ~int64_t(1);
    cc = std::min<int64_t>(cc, (ncols + 1) & ~int64_t(1));
    const int NS = 3;
    std::shared_ptr<DevBuf> xd[NS], yd[NS];
    // one contiguous copy when the host matrix is packed, else a pitched copy
    auto copy2d = [](void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows,
                     cudaMemcpyKind kind, cudaStream_t st) {
        if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, st);
        return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st);
    };
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[NS], ev_cmp[NS], ev_out[NS];
    int st = TB200_OK;
    for (int i = 0; i < NS; ++i) { ev_in[i] = ev_cmp[i] = ev_out[i] = nullptr; }
    auto cleanup = [&]() {
        cudaDeviceSynchronize();
        for (int i = 0; i < NS; ++i) {
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_cmp[i]) cudaEventDestroy(ev_cmp[i]);
            if (ev_out[i]) cudaEventDestroy(ev_out[i]);
        }
    };
#define TB_HOSTTRY(expr) do { st = (expr); if (st != TB200_OK) { cleanup(); return st; } } while (0)
    for (int i = 0; i < 3; ++i)
        if (!h->host_s[i]) TB_HOSTTRY(cuda_status(cudaStreamCreateWithFlags(&h->host_s[i], cudaStreamNonBlocking), "stream"));
    s_in = h->host_s[0]; s_cmp = h->host_s[1]; s_out = h->host_s[2];
    for (int i = 0; i < NS; ++i) {
        TB_HOSTTRY(dev_alloc_sync(xd[i], (size_t)cc * h->n * xsz, "host-path x buffer"));
        TB_HOSTTRY(dev_alloc_sync(yd[i], (size_t)cc * h->m * ysz, "host-path y buffer"));
        TB_HOSTTRY(cuda_status(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming), "event"));
        TB_HOSTTRY(cuda_status(cudaEventCreateWithFlags(&ev_cmp[i], cudaEventDisableTiming), "event"));
        TB_HOSTTRY(cuda_status(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming), "event"));
    }
    int64_t chunk = 0;
    for (int64_t c0 = 0; c0 < ncols; c0 += cc, ++chunk) {
        const int sl = (int)(chunk % NS);
        const int64_t nc = std::min(cc, ncols - c0);
        const char* xs = reinterpret_cast<const char*>(x_host) + (size_t)c0 * ldx * xsz;
        char* ys = reinterpret_cast<char*>(y_host) + (size_t)c0 * ldy * ysz;
        // H2D (x buffer free once the previous compute on this slot is done; y buffer once drained)
        if (chunk >= NS) {
            TB_HOSTTRY(cuda_status(cudaStreamWaitEvent(s_in, ev_cmp[sl], 0), "wait"));
            TB_HOSTTRY(cuda_status(cudaStreamWaitEvent(s_in, ev_out[sl], 0), "wait"));
        }
        TB_HOSTTRY(cuda_status(copy2d(xd[sl]->p, (size_t)h->n * xsz, xs, (size_t)ldx * xsz, (size_t)h->n * xsz,
                                      (size_t)nc, cudaMemcpyHostToDevice, s_in), "x upload"));
        if (has_beta)
            TB_HOSTTRY(cuda_status(copy2d(yd[sl]->p, (size_t)h->m * ysz, ys, (size_t)ldy * ysz, (size_t)h->m * ysz,
                                          (size_t)nc, cudaMemcpyHostToDevice, s_in), "y upload"));
        TB_HOSTTRY(cuda_status(cudaEventRecord(ev_in[sl], s_in), "record"));
        // compute
        TB_HOSTTRY(cuda_status(cudaStreamWaitEvent(s_cmp, ev_in[sl], 0), "wait"));
        TB_HOSTTRY(mul_impl(h, h->spec->p, false, yd[sl]->p, h->m, ydtype, xd[sl]->p, h->n, xdtype, nc, alpha, beta, s_cmp));
        TB_HOSTTRY(cuda_status(cudaEventRecord(ev_cmp[sl], s_cmp), "record"));
        // D2H
        TB_HOSTTRY(cuda_status(cudaStreamWaitEvent(s_out, ev_cmp[sl], 0), "wait"));
        TB_HOSTTRY(cuda_status(copy2d(ys, (size_t)ldy * ysz, yd[sl]->p, (size_t)h->m * ysz, (size_t)h->m * ysz,
                                      (size_t)nc, cudaMemcpyDeviceToHost, s_out), "y download"));
        TB_HOSTTRY(cuda_status(cudaEventRecord(ev_out[sl], s_out), "record"));
    }
    TB_HOSTTRY(cuda_status(cudaStreamSynchronize(s_out), "host-path sync"));
#undef TB_HOSTTRY
    cleanup();
    return TB200_OK;
}

int tb200_mul_direct(int kind, int dtype, int64_t m, int64_t n, const void* vc, const void* vr, void* y, int64_t ldy,
                     int ydtype, const void* x, int64_t ldx, int xdtype, int64_t ncols, const double alpha[2],
                     const double beta[2], void* stream) {
    TB_TRY(check_factorize_args(kind, dtype, m, n, vc, vr, true));
    if (!dt_valid(xdtype) || !dt_valid(ydtype)) return set_error(TB200_BAD_ARG, "unknown dtype");
    const int prec = dt_prec(dtype);
    if (dt_prec(xdtype) != prec || dt_prec(ydtype) != prec) return set_error(TB200_BAD_ARG, "precision mismatch");
    if (ncols <= 0) return TB200_OK;
    if (ldx < n || ldy < m) return set_error(TB200_DIM_MISMATCH, "leading dimensions smaller than the matrix size");
    const double a0 = alpha ? alpha[0] : 1.0, a1 = alpha ? alpha[1] : 0.0;
    const double b0 = beta ? beta[0] : 0.0, b1 = beta ? beta[1] : 0.0;
    const int has_beta = (b0 != 0.0 || b1 != 0.0);
    if (!dt_is_cplx(ydtype) && (dt_is_cplx(dtype) || dt_is_cplx(xdtype) || a1 != 0.0 || b1 != 0.0))
        return set_error(TB200_BAD_ARG, "a real y requires a real matrix, real x and real alpha/beta");
    dim3 grid((unsigned)((m + 127) / 128), (unsigned)ncols);
    cudaStream_t s = (cudaStream_t)stream;
    if (prec)
        k_direct_mul<double><<<grid, 128, 0, s>>>(y, ldy, dt_is_cplx(ydtype), x, ldx, dt_is_cplx(xdtype), kind, vc, vr,
                                                  dt_is_cplx(dtype), m, n, mk<double>(a0, a1), mk<double>(b0, b1), has_beta);
    else
        k_direct_mul<float><<<grid, 128, 0, s>>>(y, ldy, dt_is_cplx(ydtype), x, ldx, dt_is_cplx(xdtype), kind, vc, vr,
                                                 dt_is_cplx(dtype), m, n, mk<float>((float)a0, (float)a1),
                                                 mk<float>((float)b0, (float)b1), has_beta);
    count_launch();
    return cuda_status(cudaGetLastError(), "k_direct_mul");
}

// ---- Circulant family ------------------------------------------------------------------------
static int require_circulant(tb200_fact* h) {
    if (!h) return set_error(TB200_BAD_ARG, "handle is NULL");
    if (h->kind != TB200_CIRCULANT) return set_error(TB200_BAD_ARG, "operation is only defined for Circulant factorizations");
    return TB200_OK;
}

static int spec_map_launch(tb200_fact* h, void* out, const void* in, int op, double tol, cudaStream_t s) {
    if (h->prec) k_spec_map<double><<<ew_grid(h->Np), 256, 0, s>>>((cx<double>*)out, (const cx<double>*)in, h->Np, op, tol);
    else k_spec_map<float><<<ew_grid(h->Np), 256, 0, s>>>((cx<float>*)out, (const cx<float>*)in, h->Np, op, (float)tol);
    count_launch();
    return cuda_status(cudaGetLastError(), "k_spec_map");
}

static int ensure_inv_spec(tb200_fact* h, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(h->mu);
    if (h->inv_spec) return TB200_OK;
    std::shared_ptr<DevBuf> b;
    TB_TRY(dev_alloc_on(b, (size_t)h->Np * (h->prec ? 16 : 8), "reciprocal spectrum", s));
    TB_TRY(wait_spectrum(h, nullptr, s));
    TB_TRY(spec_map_launch(h, b->p, h->spec->p, OP_INV, 0.0, s));
    if (!h->ev_inv) TB_CUDA(cudaEventCreateWithFlags(&h->ev_inv, cudaEventDisableTiming), "event");
    h->inv_done.store(false);
    TB_CUDA(cudaEventRecord(h->ev_inv, s), "reciprocal spectrum ready");
    h->inv_stream = s;
    h->inv_spec = b;
    return TB200_OK;
}

int tb200_circ_ldiv(tb200_handle_t h, void* x, int64_t ldx, const void* b, int64_t ldb, int dtype, int64_t ncols,
                    void* stream) {
    DeviceGuard dev_guard(h);
    TB_TRY(require_circulant(h));
    TB_TRY(ensure_inv_spec(h, (cudaStream_t)stream));
    const double one[2] = {1, 0}, zero[2] = {0, 0};
    // tmp ./= vcvr_dft is applied as a multiply by the cached reciprocal spectrum
    return mul_impl(h, h->inv_spec->p, true, x, ldx, dtype, b, ldb, dtype, ncols, one, zero, (cudaStream_t)stream);
}

int tb200_circ_spectrum(tb200_handle_t h, void* out, void* stream) {
    DeviceGuard dev_guard(h);
    TB_TRY(require_circulant(h));
    if (!out) return set_error(TB200_BAD_ARG, "out is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    TB_TRY(wait_spectrum(h, nullptr, s));
    if (h->bluestein)
        return cuda_status(cudaMemcpyAsync(out, h->spec->p, (size_t)h->n * (h->prec ? 16 : 8), cudaMemcpyDeviceToDevice, s), "spectrum copy");
    // three-level: outer digit k0 (2^log2l1 points), then the child plan's two levels inside each row
    const int l0 = h->three_level ? h->log2l1 : 0;
    const tb200_fact* pl = h->three_level ? h->inner.get() : h;
    const int l1 = pl->two_level ? pl->log2l1 : 0;
    if (h->prec) k_unscramble<double><<<ew_grid(h->Np), 256, 0, s>>>((cx<double>*)out, (const cx<double>*)h->spec->p, h->Np, l0, l1, pl->log2l2, le_of<double>());
    else k_unscramble<float><<<ew_grid(h->Np), 256, 0, s>>>((cx<float>*)out, (const cx<float>*)h->spec->p, h->Np, l0, l1, pl->log2l2, le_of<float>());
    count_launch();
    return cuda_status(cudaGetLastError(), "k_unscramble");
}

static int clone_plan(tb200_fact* h, cudaStream_t s, tb200_fact** out) {
    tb200_fact* g = nullptr;
    TB_TRY(new_handle(h->kind, h->dtype, h->m, h->n, s, &g));
    if (h->bluestein) { g->bs_conv = h->bs_conv; g->bs_chirp = h->bs_chirp; }
    *out = g;
    return TB200_OK;
}

int tb200_circ_map(tb200_handle_t h, int op, double tol, void* stream, tb200_handle_t* out) {
    DeviceGuard dev_guard(h);
    TB_TRY(require_circulant(h));
    if (!out) return set_error(TB200_BAD_ARG, "out is NULL");
    if (op < 0 || op > 4) return set_error(TB200_BAD_ARG, "unknown spectrum op %d", op);
    tb200_fact* g = nullptr;
    TB_TRY(clone_plan(h, (cudaStream_t)stream, &g));
    int st = wait_spectrum(h, nullptr, (cudaStream_t)stream);
    if (st == TB200_OK) st = spec_map_launch(h, g->spec->p, h->spec->p, op, tol, (cudaStream_t)stream);
    if (st == TB200_OK) st = mark_spectrum_ready(g, (cudaStream_t)stream);
    if (st != TB200_OK) { delete g; return st; }
    // sqrt/inv/pinv/conj of a real circulant's spectrum stay Hermitian only for inv/pinv/conj; sqrt of a
    // spectrum with negative real eigenvalues does not, but the reference takes maybereal anyway (:314).
    *out = g;
    return TB200_OK;
}

int tb200_circ_mul(tb200_handle_t a, tb200_handle_t b, void* stream, tb200_handle_t* out) {
    DeviceGuard dev_guard(a);
    TB_TRY(require_circulant(a));
    TB_TRY(require_circulant(b));
    if (!out) return set_error(TB200_BAD_ARG, "out is NULL");
    if (a->n != b->n)
        return set_error(TB200_DIM_MISMATCH, "size of matrix A, %lldx%lld, does not match size of matrix B, %lldx%lld",
                         (long long)a->n, (long long)a->n, (long long)b->n, (long long)b->n);
    if (a->prec != b->prec) return set_error(TB200_BAD_ARG, "precision mismatch between the two factorizations");
    tb200_fact* g = nullptr;
    const int dt = (dt_is_cplx(a->dtype) || dt_is_cplx(b->dtype)) ? (a->prec ? TB200_C64 : TB200_C32) : a->dtype;
    cudaStream_t s = (cudaStream_t)stream;
    TB_TRY(new_handle(TB200_CIRCULANT, dt, a->n, a->n, s, &g));
    if (a->bluestein) { g->bs_conv = a->bs_conv; g->bs_chirp = a->bs_chirp; }
    TB_TRY(wait_spectrum(a, nullptr, s));
    TB_TRY(wait_spectrum(b, nullptr, s));
    if (a->prec) k_spec_mul<double><<<ew_grid(a->Np), 256, 0, s>>>((cx<double>*)g->spec->p, (const cx<double>*)a->spec->p, (const cx<double>*)b->spec->p, a->Np);
    else k_spec_mul<float><<<ew_grid(a->Np), 256, 0, s>>>((cx<float>*)g->spec->p, (const cx<float>*)a->spec->p, (const cx<float>*)b->spec->p, a->Np);
    count_launch();
    int st = cuda_status(cudaGetLastError(), "k_spec_mul");
    if (st == TB200_OK) st = mark_spectrum_ready(g, s);
    if (st != TB200_OK) { delete g; return st; }
    *out = g;
    return TB200_OK;
}

int tb200_circ_to_vc(tb200_handle_t h, void* vc_out, int out_dtype, void* stream) {
    DeviceGuard dev_guard(h);
    TB_TRY(require_circulant(h));
    if (!vc_out) return set_error(TB200_BAD_ARG, "vc_out is NULL");
    if (!dt_valid(out_dtype) || dt_prec(out_dtype) != h->prec) return set_error(TB200_BAD_ARG, "output dtype must have the factorization's precision");
    if (h->bluestein)
        return bs_run(h, 1, nullptr, vc_out, h->n, dt_is_cplx(out_dtype), 0, h->spec->p, h->n, 1, 1, nullptr, nullptr, (cudaStream_t)stream);
    ConvCall c;
    c.x = h->spec->p; c.ldx = h->Np; c.x_mode = IO_SPEC; c.n_in = h->Np;
    c.y = vc_out; c.ldy = h->Np; c.y_mode = dt_is_cplx(out_dtype) ? IO_CPLX : IO_REAL; c.m_out = h->n;
    c.ncols = 1; c.nfft = 1; c.do_fwd = 0; c.do_inv = 1; c.spec = nullptr;
    c.scale = 1.0 / (double)h->Np;
    return run_conv(h, c, (cudaStream_t)stream);
}

int tb200_precond_vector(int which, int kind, int dtype, int64_t n, const void* vc, const void* vr, void* v_out,
                         void* stream) {
    TB_TRY(check_factorize_args(kind, dtype, n, n, vc, vr, true));
    if (kind == TB200_HANKEL) return set_error(TB200_BAD_ARG, "strang/chan are defined for Toeplitz-structured matrices");
    if (which != 0 && which != 1) return set_error(TB200_BAD_ARG, "which must be 0 (strang) or 1 (chan)");
    if (!v_out) return set_error(TB200_BAD_ARG, "v_out is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (dt_prec(dtype)) k_strang_chan<double><<<ew_grid(n), 256, 0, s>>>(v_out, which, kind, vc, vr, dt_is_cplx(dtype), n);
    else k_strang_chan<float><<<ew_grid(n), 256, 0, s>>>(v_out, which, kind, vc, vr, dt_is_cplx(dtype), n);
    count_launch();
    return cuda_status(cudaGetLastError(), "k_strang_chan");
}

// ---- Levinson / Durbin -----------------------------------------------------------------------
int tb200_levinson_batched(int dtype, int64_t n, int64_t nsys, const void* r, int64_t ldr, const void* b, int64_t ldb,
                           void* x, int64_t ldx, const void* diag, void* stream) {
    if (n <= 0 || nsys < 0) return set_error(TB200_BAD_ARG, "bad sizes");
    if (nsys == 0) return TB200_OK;
    if (!b || !x || (n > 1 && !r)) return set_error(TB200_BAD_ARG, "NULL argument");
    if (ldb < n || ldx < n || (n > 1 && ldr < n - 1))
        return set_error(TB200_DIM_MISMATCH, "length(x) = %lld / length(b) = %lld != %lld = length(r) + 1", (long long)ldx, (long long)ldb, (long long)n);
    return levinson_dispatch(dtype, false, n, nsys, r, ldr, b, ldb, x, ldx, diag, (cudaStream_t)stream);
}

int tb200_durbin_batched(int dtype, int64_t n, int64_t nsys, const void* r, int64_t ldr, void* y, int64_t ldy, void* stream) {
    if (n <= 0 || nsys < 0) return set_error(TB200_BAD_ARG, "bad sizes");
    if (nsys == 0) return TB200_OK;
    if (!r || !y) return set_error(TB200_BAD_ARG, "NULL argument");
    if (ldr < n || ldy < n) return set_error(TB200_DIM_MISMATCH, "length(y) = %lld != %lld = length(r)", (long long)ldy, (long long)n);
    return levinson_dispatch(dtype, true, n, nsys, r, ldr, nullptr, 0, y, ldy, nullptr, (cudaStream_t)stream);
}

int tb200_levinson_multirhs(int dtype, int64_t n, int64_t nrhs, const void* vc, const void* B, int64_t ldb, void* X, int64_t ldx,
                            void* stream) {
    if (n <= 0 || nrhs < 0) return set_error(TB200_BAD_ARG, "bad sizes");
    if (nrhs == 0) return TB200_OK;
    if (!vc || !B || !X) return set_error(TB200_BAD_ARG, "NULL argument");
    if (ldb < n || ldx < n) return set_error(TB200_DIM_MISMATCH, "B and X must have %lld rows", (long long)n);
    if (dtype != TB200_F32 && dtype != TB200_F64) return set_error(TB200_UNSUPPORTED, "levinson: only real Float32/Float64 systems");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t esz = dt_prec(dtype) ? 8 : 4;
    std::shared_ptr<DevBuf> work;
    TB_TRY(dev_alloc_on(work, esz * (size_t)(n * ((n + 127) / 128 * 128) + 2 * n + 8), "levinson y-table", s));
    TB_TRY(levinson_multirhs_dispatch(dtype, n, nrhs, vc, B, ldb, X, ldx, work->p, s));
    TB_CUDA(cudaStreamSynchronize(s), "levinson multi-RHS sync");     // the table is released on return
    return TB200_OK;
}

int tb200_vec_op(int op, int64_t n, const void* a, int dtype_a, int rev_a, const void* b, int dtype_b, int rev_b, void* out,
                 int dtype_out, const double alpha[2], const double beta[2], void* stream) {
    if (op < 0 || op > VOP_FILL) return set_error(TB200_BAD_ARG, "unknown vector op %d", op);
    if (n < 0) return set_error(TB200_BAD_ARG, "bad length");
    if (n == 0) return TB200_OK;
    if (!out || (!a && op != VOP_FILL)) return set_error(TB200_BAD_ARG, "NULL argument");
    if (!dt_valid(dtype_out) || (a && !dt_valid(dtype_a)) || (b && !dt_valid(dtype_b))) return set_error(TB200_BAD_ARG, "unknown dtype");
    const double2 al = make_double2(alpha ? alpha[0] : 1.0, alpha ? alpha[1] : 0.0);
    const double2 be = make_double2(beta ? beta[0] : 0.0, beta ? beta[1] : 0.0);
    k_vec_op<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(op, n, a, dtype_a, rev_a, b, dtype_b, rev_b, out, dtype_out, al, be);
    count_launch();
    return cuda_status(cudaGetLastError(), "k_vec_op");
}

int tb200_vec_equal(int64_t n, const void* a, int dtype_a, const void* b, int dtype_b, int* equal, void* stream) {
    if (!equal) return set_error(TB200_BAD_ARG, "NULL argument");
    *equal = 1;
    if (n <= 0) return TB200_OK;
    if (!a || !dt_valid(dtype_a) || (b && !dt_valid(dtype_b))) return set_error(TB200_BAD_ARG, "bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    std::shared_ptr<DevBuf> cnt;
    TB_TRY(dev_alloc_on(cnt, sizeof(unsigned long long), "mismatch counter", s));
    TB_CUDA(cudaMemsetAsync(cnt->p, 0, sizeof(unsigned long long), s), "memset");
    k_vec_mismatch<<<ew_grid(n), 256, 0, s>>>(n, a, dtype_a, b, dtype_b, (unsigned long long*)cnt->p);
    count_launch();
    TB_CUDA(cudaGetLastError(), "k_vec_mismatch");
    unsigned long long hcnt = 0;
    TB_CUDA(cudaMemcpyAsync(&hcnt, cnt->p, sizeof(hcnt), cudaMemcpyDeviceToHost, s), "mismatch count");
    TB_CUDA(cudaStreamSynchronize(s), "mismatch sync");
    TB_TRY(retire_after(s));
    *equal = hcnt == 0 ? 1 : 0;
    return TB200_OK;
}

int tb200_tri_smallinv(int dtype, int64_t n, const void* v, void* out, void* stream) {
    if (!dt_valid(dtype)) return set_error(TB200_BAD_ARG, "unknown dtype");
    if (n < 1 || n > 64) return set_error(TB200_BAD_ARG, "smallinv is the n <= 64 base case, got n = %lld", (long long)n);
    if (!v || !out) return set_error(TB200_BAD_ARG, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (dt_prec(dtype)) k_tri_smallinv<double><<<1, 32, 0, s>>>(out, v, dt_is_cplx(dtype), (int)n);
    else k_tri_smallinv<float><<<1, 32, 0, s>>>(out, v, dt_is_cplx(dtype), (int)n);
    count_launch();
    return cuda_status(cudaGetLastError(), "k_tri_smallinv");
}

int tb200_trench(int dtype, int64_t n, const void* r, void* B, int64_t ldb, double diag, void* stream) {
    if (dtype != TB200_F32 && dtype != TB200_F64) return set_error(TB200_UNSUPPORTED, "trench: real Float32/Float64 only");
    if (n < 1) return set_error(TB200_BAD_ARG, "bad size");
    if (!B || (n > 1 && !r)) return set_error(TB200_BAD_ARG, "NULL argument");
    if (ldb < n) return set_error(TB200_DIM_MISMATCH, "B must be %lld x %lld", (long long)n, (long long)n);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t esz = dt_prec(dtype) ? 8 : 4;
    std::shared_ptr<DevBuf> ybuf, scal;
    TB_TRY(dev_alloc_on(ybuf, esz * (size_t)std::max<int64_t>(n - 1, 1), "trench durbin vector", s));
    TB_TRY(dev_alloc_on(scal, sizeof(double) * 4 * 64 + 64, "trench scalars", s));
    double* dots = (double*)scal->p;
    double* partials = dots + 4;
    unsigned* counter = (unsigned*)(partials + 4 * 60);
    TB_CUDA(cudaMemsetAsync(scal->p, 0, scal->bytes, s), "memset");
    if (n > 1) {
        TB_TRY(levinson_dispatch(dtype, true, n - 1, 1, r, n - 1, nullptr, 0, ybuf->p, n - 1, nullptr, s));   // y = durbin(r)
        const unsigned grid = std::min<unsigned>(ew_grid(n - 1), 60u);
        if (dt_prec(dtype)) k_dot2<double><<<grid, 256, 0, s>>>((const double*)r, (const double*)ybuf->p, nullptr, nullptr, n - 1, partials, counter, dots);
        else k_dot2<float><<<grid, 256, 0, s>>>((const float*)r, (const float*)ybuf->p, nullptr, nullptr, n - 1, partials, counter, dots);
        count_launch();
    }
    const unsigned g2 = (unsigned)((n + 127) / 128);
    if (dt_prec(dtype)) k_trench_fill<double><<<g2, 128, 0, s>>>((double*)B, ldb, (const double*)ybuf->p, dots, n, 1.0 / diag);
    else k_trench_fill<float><<<g2, 128, 0, s>>>((float*)B, ldb, (const float*)ybuf->p, dots, n, (float)(1.0 / diag));
    count_launch();
    TB_CUDA(cudaGetLastError(), "trench kernels");
    TB_CUDA(cudaStreamSynchronize(s), "trench sync");       // temporaries are released on return
    return TB200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// iterative solvers
// ---------------------------------------------------------------------------------------
namespace tb {

static int check_krylov(tb200_fact* A, tb200_fact* M, int dtype, int64_t n, const void* x, const void* b) {
    if (!A) return set_error(TB200_BAD_ARG, "A is NULL");
    if (!x || !b) return set_error(TB200_BAD_ARG, "x or b is NULL");
    if (!dt_valid(dtype) || dt_prec(dtype) != A->prec) return set_error(TB200_BAD_ARG, "vector precision must match A");
    if (A->m != A->n || A->n != n) return set_error(TB200_DIM_MISMATCH, "A is %lldx%lld but the vectors have length %lld", (long long)A->m, (long long)A->n, (long long)n);
    if (!dt_is_cplx(dtype) && !A->a_real) return set_error(TB200_BAD_ARG, "real vectors with a complex matrix");
    if (M) {
        if (M->kind != TB200_CIRCULANT || M->n != n) return set_error(TB200_DIM_MISMATCH, "preconditioner must be a Circulant factorization of size %lld", (long long)n);
        if (M->prec != A->prec) return set_error(TB200_BAD_ARG, "preconditioner precision mismatch");
        if (!dt_is_cplx(dtype) && !M->a_real) return set_error(TB200_BAD_ARG, "real vectors with a complex preconditioner");
    }
    return TB200_OK;
}

// IterativeLinearSolvers.cgs (src/iterativeLinearSolvers.jl:99-215) and cg (:9-97) over the columns of B; one column
// is the single-vector solver.  krylov_batched.cuh has the kernels and the per-column state machine.
template <typename T>
static int krylov_batched_t(bool is_cg, tb200_fact* A, tb200_fact* M, T* X, int64_t ldx, const T* B, int64_t ldb, int dtype,
                            int64_t n, int64_t nrhs, int max_it, double tol, double* err, int* iter, int* flag, cudaStream_t s) {
    const int nvec = is_cg ? 4 : 8;
    int64_t chunk = std::max<int64_t>(1, g_krylov_work_bytes.load() / (int64_t)(sizeof(T) * (size_t)n * nvec));
    chunk = std::min<int64_t>(std::min<int64_t>(chunk, nrhs), 65535);
    if (chunk > 1 && (chunk & 1)) --chunk;                       // keep real columns paired in the transforms
    std::shared_ptr<DevBuf> work, stbuf, partials, counters, nact, cmap;
    TB_TRY(dev_alloc_on(work, sizeof(T) * (size_t)n * (size_t)chunk * nvec, "batched Krylov work vectors", s));
    // CTAs per column of the dot-product sweeps: enough to fill the chip when there are few columns
    const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>(std::max<long long>(128, 2368 / chunk), (n + 2047) / 2048));
    TB_TRY(dev_alloc_on(stbuf, sizeof(KState) * (size_t)chunk, "Krylov state", s));
    TB_TRY(dev_alloc_on(partials, sizeof(double) * 4 * gx * (size_t)chunk, "dot partials", s));
    TB_TRY(dev_alloc_on(counters, sizeof(unsigned) * (size_t)chunk, "dot counters", s));
    TB_TRY(dev_alloc_on(nact, sizeof(int), "active counter", s));
    TB_TRY(dev_alloc_on(cmap, sizeof(int) * (size_t)chunk, "active column map", s));
    KState* st = (KState*)stbuf->p;
    int* colmap = (int*)cmap->p;
    std::vector<KState> host_st((size_t)chunk);
    const double one[2] = {1, 0}, zero[2] = {0, 0}, minus[2] = {-1, 0};
    const size_t vsz = (size_t)n * (size_t)chunk;
    T* w = (T*)work->p;
    // column selection needs the pow2 kernels on both operators (a Bluestein plan runs every column)
    // column selection / per-column alpha live in the two-level passes; chirp-z and three-level plans take the unfused route
    const bool can_select = !A->bluestein && !A->three_level && !(M && (M->bluestein || M->three_level));
    auto finish = [&](int status) {                     // temporaries are stream-ordered: drain before they go
        cudaStreamSynchronize(s);
        return status;
    };
#define TB_KTRY(expr) do { int st__ = (expr); if (st__ != TB200_OK) return finish(st__); } while (0)
    for (int64_t c0 = 0; c0 < nrhs; c0 += chunk) {
        const int64_t c = std::min(chunk, nrhs - c0);
        T* Xc = X + c0 * ldx;
        const T* Bc = B + c0 * ldb;
        const unsigned gxe = (unsigned)std::max<long long>(1, std::min<long long>((n + 1023) / 1024, 2048));
        const unsigned gs = (unsigned)((c + 127) / 128);
        int h_active = 0;            // columns still iterating, as last read from the device (never below the true count)
        const int* cm = nullptr;     // column map handed to the kernels (NULL while every column is live)
        auto dots = [&](const T* a1, const T* b1, long long ld1, const T* a2, const T* b2, int only_active, int ncol) -> int {
            k_bdot2<T><<<dim3(gx, (unsigned)ncol), 256, 0, s>>>(a1, b1, ld1, a2, b2, n, n, (double*)partials->p, (unsigned*)counters->p,
                                                              st, only_active, only_active ? cm : nullptr);
            count_launch();
            return cuda_status(cudaGetLastError(), "k_bdot2");
        };
        auto step = [&](int which, int it) -> int {
            k_kstep<<<gs, 128, 0, s>>>(st, (int)c, which, it, max_it, tol);
            count_launch();
            return cuda_status(cudaGetLastError(), "k_kstep");
        };
        // partition the columns and (optionally) read the live count back
        auto compact = [&](bool readback) -> int {
            k_compact<<<1, 1024, 0, s>>>(st, (int)c, colmap, (int*)nact->p);
            count_launch();
            TB_CUDA(cudaGetLastError(), "k_compact");
            if (readback) {
                TB_CUDA(cudaMemcpyAsync(&h_active, nact->p, sizeof(int), cudaMemcpyDeviceToHost, s), "active readback");
                TB_CUDA(cudaStreamSynchronize(s), "Krylov sync");
                cm = (can_select && h_active < c) ? colmap : nullptr;
            }
            return TB200_OK;
        };
        auto mulA = [&](T* y, const T* x, int64_t ldxx, int64_t ncol, const double* a, const double* b, bool all, bool per_col_alpha) -> int {
            ColSel sel;
            if (!all) sel.colmap = cm;
            if (per_col_alpha) { sel.col_alpha = &st[0].alpha[0]; sel.col_alpha_stride = (int)(sizeof(KState) / sizeof(double)); }
            return mul_impl(A, A->spec->p, false, y, n, dtype, x, ldxx, dtype, (all || !cm) ? c : ncol, a, b, s, &sel);
        };
        auto precond = [&](T* out_v, const T* in_v, int64_t ncol) -> int {
            if (!M) {
                if (out_v != in_v) TB_CUDA(cudaMemcpyAsync(out_v, in_v, sizeof(T) * (size_t)n * (size_t)c, cudaMemcpyDeviceToDevice, s), "copy");
                return TB200_OK;
            }
            TB_TRY(ensure_inv_spec(M, s));
            ColSel sel;
            sel.colmap = cm;
            return mul_impl(M, M->inv_spec->p, true, out_v, n, dtype, in_v, n, dtype, cm ? ncol : c, one, zero, s, &sel);
        };
        // x_j and r_j = b_j enter the iteration scaled by the column's power of two (k_kstep KS_BNRM)
        auto scaled_start = [&](T* r) -> int {
            k_bscale<T><<<dim3(gxe, (unsigned)c), 256, 0, s>>>(Xc, ldx, st, n, 0);
            count_launch();
            k_bcopy_scaled<T><<<dim3(gxe, (unsigned)c), 256, 0, s>>>(r, n, Bc, ldb, st, n);
            count_launch();
            return cuda_status(cudaGetLastError(), "k_bcopy_scaled");
        };
        // small problems are launch-bound: read the live count every third iteration only (stale counts are safe, see k_compact)
        auto sync_now = [&](int it) { return (long long)n * std::max(h_active, 1) >= (1ll << 19) || it % 3 == 0 || it == max_it; };
        TB_KTRY(cuda_status(cudaMemsetAsync(counters->p, 0, sizeof(unsigned) * (size_t)c, s), "memset"));
        TB_KTRY(cuda_status(cudaMemsetAsync(work->p, 0, sizeof(T) * vsz * nvec, s), "memset"));          // :141-146 zeros
        TB_KTRY(dots(Bc, Bc, ldb, nullptr, nullptr, 0, (int)c));
        TB_KTRY(step(KS_BNRM, (std::is_same<T, float>::value || std::is_same<T, float2>::value) ? 100 : 900));
        if (!is_cg) {
            T* u = w; T* p = u + vsz; T* ph = p + vsz; T* q = ph + vsz; T* uh = q + vsz; T* vh = uh + vsz; T* r = vh + vsz; T* rt = r + vsz;
            TB_KTRY(scaled_start(r));                                                                       // :149 (in units of ~||b_j||)
            TB_KTRY(mulA(r, Xc, ldx, c, minus, one, true, false));                                          // :150
            TB_KTRY(dots(r, r, n, nullptr, nullptr, 0, (int)c));
            TB_KTRY(step(KS_INIT_CGS, 0));
            TB_KTRY(compact(true));
            if (h_active > 0) {
                TB_KTRY(cuda_status(cudaMemcpyAsync(rt, r, sizeof(T) * (size_t)n * (size_t)c, cudaMemcpyDeviceToDevice, s), "copy"));   // :158
                TB_KTRY(dots(rt, r, n, nullptr, nullptr, 1, cm ? h_active : (int)c));
                TB_KTRY(step(KS_RHO0, 0));
                for (int it = 1; it <= max_it && h_active > 0; ++it) {
                    const int na = cm ? h_active : (int)c;
                    const dim3 gew(gxe, (unsigned)na);
                    TB_KTRY(step(KS_CGS_BEGIN, it));
                    k_bcgs_dir<T><<<gew, 256, 0, s>>>(u, p, r, q, st, it == 1, n, n, cm);                   // :168-177
                    count_launch();
                    TB_KTRY(precond(ph, p, na));                                                            // :179-180
                    TB_KTRY(mulA(vh, ph, n, na, one, zero, false, false));                                  // :182
                    TB_KTRY(dots(rt, vh, n, nullptr, nullptr, 1, na));
                    TB_KTRY(step(KS_CGS_ALPHA, it));                                                        // :184
                    k_bcgs_q<T><<<gew, 256, 0, s>>>(q, uh, u, vh, st, n, n, cm);                            // :185-188
                    count_launch();
                    TB_KTRY(precond(uh, uh, na));                                                           // :189
                    if (!A->bluestein && !A->three_level) {
                        k_bx<T><<<gew, 256, 0, s>>>(Xc, ldx, uh, st, n, n, cm);                             // :192-194
                        count_launch();
                        TB_KTRY(mulA(r, uh, n, na, minus, one, false, true));                               // :196  r -= alpha_j * A * uh
                    } else {                                                                                // chirp-z plan: unfused
                        TB_KTRY(mulA(vh, uh, n, na, one, zero, false, false));
                        k_bxr<T><<<gew, 256, 0, s>>>(Xc, ldx, r, uh, vh, st, n, n, cm);
                        count_launch();
                    }
                    TB_KTRY(dots(r, r, n, rt, r, 1, na));
                    TB_KTRY(step(KS_CGS_END, it));                                                          // :198-201
                    TB_KTRY(compact(sync_now(it)));
                }
            }
        } else {
            T* z = w; T* q = z + vsz; T* p = q + vsz; T* r = p + vsz;
            TB_KTRY(scaled_start(r));
            TB_KTRY(mulA(r, Xc, ldx, c, minus, one, true, false));                                          // :55
            TB_KTRY(dots(r, r, n, nullptr, nullptr, 0, (int)c));
            TB_KTRY(step(KS_INIT_CG, 0));
            TB_KTRY(compact(true));
            for (int it = 1; it <= max_it && h_active > 0; ++it) {
                const int na = cm ? h_active : (int)c;
                const dim3 gew(gxe, (unsigned)na);
                TB_KTRY(precond(z, r, na));                                                                 // :64-65
                TB_KTRY(dots(r, z, n, nullptr, nullptr, 1, na));
                TB_KTRY(step(KS_CG_BEGIN, it));                                                             // :67-68
                k_bcg_dir<T><<<gew, 256, 0, s>>>(p, z, st, it == 1, n, n, cm);                              // :69-76
                count_launch();
                TB_KTRY(mulA(q, p, n, na, one, zero, false, false));                                        // :79
                TB_KTRY(dots(p, q, n, nullptr, nullptr, 1, na));
                TB_KTRY(step(KS_CG_ALPHA, it));                                                             // :80
                k_bxr<T><<<gew, 256, 0, s>>>(Xc, ldx, r, p, q, st, n, n, cm);                               // :81-84
                count_launch();
                TB_KTRY(dots(r, r, n, nullptr, nullptr, 1, na));
                TB_KTRY(step(KS_CG_END, it));                                                               // :86-88
                TB_KTRY(compact(sync_now(it)));
            }
        }
        TB_KTRY(cuda_status(cudaGetLastError(), "batched Krylov kernels"));
        k_bscale<T><<<dim3(gxe, (unsigned)c), 256, 0, s>>>(Xc, ldx, st, n, 1);                             // back to the units of b_j
        count_launch();
        TB_KTRY(cuda_status(cudaMemcpyAsync(host_st.data(), st, sizeof(KState) * (size_t)c, cudaMemcpyDeviceToHost, s), "state readback"));
        TB_KTRY(cuda_status(cudaStreamSynchronize(s), "Krylov sync"));
        bool any_nan = false;
        for (int64_t j = 0; j < c; ++j) any_nan = any_nan || host_st[(size_t)j].pad != 0;
        if (any_nan) {
            k_bnanfill<T><<<dim3(gxe, (unsigned)c), 256, 0, s>>>(Xc, ldx, st, n);
            count_launch();
            TB_KTRY(cuda_status(cudaGetLastError(), "k_bnanfill"));
        }
        for (int64_t j = 0; j < c; ++j) {
            if (err) err[c0 + j] = host_st[(size_t)j].error;
            if (iter) iter[c0 + j] = host_st[(size_t)j].iter;
            if (flag) flag[c0 + j] = host_st[(size_t)j].flag;
        }
    }
#undef TB_KTRY
    return finish(TB200_OK);
}

static int krylov_batched(bool is_cg, tb200_fact* A, tb200_fact* M, void* X, int64_t ldx, const void* B, int64_t ldb, int dtype,
                          int64_t n, int64_t nrhs, int max_it, double tol, double* err, int* iter, int* flag, void* stream) {
    DeviceGuard dev_guard(A);
    if (nrhs < 0) return set_error(TB200_BAD_ARG, "negative column count");
    TB_TRY(check_krylov(A, M, dtype, n, X, B));
    if (ldx < n || ldb < n) return set_error(TB200_DIM_MISMATCH, "leading dimensions must be at least n = %lld", (long long)n);
    if (is_cg && dt_is_cplx(dtype)) return set_error(TB200_BAD_ARG, "cg is defined for real (BlasReal) element types only");
    if (nrhs == 0) return TB200_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
    case TB200_F32: return krylov_batched_t<float>(is_cg, A, M, (float*)X, ldx, (const float*)B, ldb, dtype, n, nrhs, max_it, tol, err, iter, flag, s);
    case TB200_F64: return krylov_batched_t<double>(is_cg, A, M, (double*)X, ldx, (const double*)B, ldb, dtype, n, nrhs, max_it, tol, err, iter, flag, s);
    case TB200_C32: return krylov_batched_t<float2>(is_cg, A, M, (float2*)X, ldx, (const float2*)B, ldb, dtype, n, nrhs, max_it, tol, err, iter, flag, s);
    default: return krylov_batched_t<double2>(is_cg, A, M, (double2*)X, ldx, (const double2*)B, ldb, dtype, n, nrhs, max_it, tol, err, iter, flag, s);
    }
}

}  // namespace tb

extern "C" {

int tb200_cgs_batched(tb200_handle_t A, tb200_handle_t M, void* X, int64_t ldx, const void* B, int64_t ldb, int dtype, int64_t n,
                      int64_t nrhs, int max_it, double tol, double* err, int* iter, int* flag, void* stream) {
    return krylov_batched(false, A, M, X, ldx, B, ldb, dtype, n, nrhs, max_it, tol, err, iter, flag, stream);
}
int tb200_cg_batched(tb200_handle_t A, tb200_handle_t M, void* X, int64_t ldx, const void* B, int64_t ldb, int dtype, int64_t n,
                     int64_t nrhs, int max_it, double tol, double* err, int* iter, int* flag, void* stream) {
    return krylov_batched(true, A, M, X, ldx, B, ldb, dtype, n, nrhs, max_it, tol, err, iter, flag, stream);
}

// single-vector solvers = the one-column case of the batched drivers (scalars stay on the device)
int tb200_cgs(tb200_handle_t A, tb200_handle_t M, void* x, const void* b, int dtype, int64_t n, int max_it, double tol,
              double* err, int* iter, int* flag, void* stream) {
    return krylov_batched(false, A, M, x, n, b, n, dtype, n, 1, max_it, tol, err, iter, flag, stream);
}

int tb200_cg(tb200_handle_t A, tb200_handle_t M, void* x, const void* b, int dtype, int64_t n, int max_it, double tol,
             double* err, int* iter, int* flag, void* stream) {
    return krylov_batched(true, A, M, x, n, b, n, dtype, n, 1, max_it, tol, err, iter, flag, stream);
}

}  // extern "C"
