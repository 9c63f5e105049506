// simetuc_b200 - C ABI implementation (host side) for include/simetuc_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "stb_build.cuh"
#include "stb_integrate.cuh"
#include "stb_integrate_cl.cuh"
#include "stb_integrate_cl2.cuh"
#include "stb_kernels.cuh"
#include "stb_tma.cuh"

static thread_local std::string g_err;
// see simetuc_b200/__init__.py: one hardware work queue per stream in flight, if nobody chose otherwise
static const int g_env_once = setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
static std::atomic<long long> g_launches{0};   // kernels launched by this library (all handles)

// ---- device memory pool ----------------------------------------------------------------------
// cudaMalloc and cudaFree wait for every kernel running on the device (measured: a 256-byte
// cudaMalloc takes 320-450 ms while an integrator kernel of another handle runs), so they have no
// place on a path where several handles are in flight.  Two layers:
//   * freed blocks are parked per device and size class and handed out again (a sweep creates and
//     destroys one system per call with the same or similar buffer sizes); at most STB_POOL_CAP
//     bytes stay parked (stb_release_cached_memory returns them to the driver);
//   * a miss is served from an arena reserved once per device before the first kernel of the
//     library runs (STB_ARENA_MB, default 1024 MB, at most 1/8 of the free memory): a first-fit
//     free list with coalescing, no driver call at all;
//   * what does not fit there goes to the stream-ordered allocator (cudaMallocAsync on a private
//     stream, release threshold "never"; measured 1-150 ms while kernels run), else cudaMalloc.
#include <mutex>
#include <unordered_map>
#define STB_POOL_CAP (8ull << 30)
#include <map>
struct PoolBlock { void* p; size_t bytes; int dev; bool async; bool arena; };
struct Arena {
    char* base = nullptr;
    size_t size = 0;
    bool tried = false;
    std::map<size_t, size_t> free_at;       // offset -> length of every free range
};
static std::mutex g_pool_mu;
static std::vector<PoolBlock> g_pool_parked;
static std::unordered_map<void*, PoolBlock> g_pool_live;
static size_t g_pool_parked_bytes = 0;
#define STB_MAX_DEV 64
static Arena g_arena[STB_MAX_DEV];

// (g_pool_mu held)  Reserve the arena of `dev` on first use.
static void arena_init(int dev) {
    Arena& A = g_arena[dev];
    if (A.tried) return;
    A.tried = true;
    size_t want = 1024;
    if (const char* env = getenv("STB_ARENA_MB")) want = (size_t)std::max(0, atoi(env));
    want <<= 20;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) want = std::min(want, free_b / 8);
    if (want < (1u << 20)) return;
    void* p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return; }
    A.base = (char*)p;
    A.size = want;
    A.free_at[0] = want;
}

// (g_pool_mu held)  First fit; bytes is a multiple of 256.
static void* arena_alloc(int dev, size_t bytes) {
    Arena& A = g_arena[dev];
    for (auto it = A.free_at.begin(); it != A.free_at.end(); ++it) {
        if (it->second < bytes) continue;
        const size_t off = it->first, len = it->second;
        A.free_at.erase(it);
        if (len > bytes) A.free_at[off + bytes] = len - bytes;
        return A.base + off;
    }
    return nullptr;
}

// (g_pool_mu held)
static void arena_free(int dev, void* p, size_t bytes) {
    Arena& A = g_arena[dev];
    size_t off = (size_t)((char*)p - A.base), len = bytes;
    auto nxt = A.free_at.lower_bound(off);
    if (nxt != A.free_at.begin()) {
        auto prv = std::prev(nxt);
        if (prv->first + prv->second == off) { off = prv->first; len += prv->second; A.free_at.erase(prv); }
    }
    if (nxt != A.free_at.end() && off + len == nxt->first) { len += nxt->second; A.free_at.erase(nxt); }
    A.free_at[off] = len;
}
static cudaStream_t g_alloc_stream[STB_MAX_DEV] = {};
static int g_alloc_state[STB_MAX_DEV] = {};      // 0: not probed, 1: stream-ordered allocator, -1: cudaMalloc only

// stream of the stream-ordered allocator on `dev`, or null when the device / driver has none
static cudaStream_t alloc_stream(int dev) {
    if (dev < 0 || dev >= STB_MAX_DEV) return nullptr;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_alloc_state[dev] == 0) {
        g_alloc_state[dev] = -1;
        int ok = 0;
        cudaMemPool_t mp;
        if (!getenv("STB_NO_MALLOC_ASYNC") &&
            cudaDeviceGetAttribute(&ok, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess && ok &&
            cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) {
            unsigned long long never = ~0ull;
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &never);
            if (cudaStreamCreateWithFlags(&g_alloc_stream[dev], cudaStreamNonBlocking) == cudaSuccess)
                g_alloc_state[dev] = 1;
        }
        cudaGetLastError();
    }
    return g_alloc_state[dev] == 1 ? g_alloc_stream[dev] : nullptr;
}

static cudaError_t raw_malloc(void** p, size_t bytes, int dev, bool* async) {
    cudaStream_t st = alloc_stream(dev);
    if (st) {
        cudaError_t e = cudaMallocAsync(p, bytes, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) { *async = true; return e; }
        cudaGetLastError();
    }
    *async = false;
    return cudaMalloc(p, bytes);
}

static void raw_free(const PoolBlock& b) {
    cudaStream_t st = b.async ? alloc_stream(b.dev) : nullptr;
    if (st) {
        cudaFreeAsync(b.p, st);
        cudaStreamSynchronize(st);
    } else {
        cudaFree(b.p);
    }
}

static cudaError_t pool_malloc(void** out, size_t bytes) {
    // size classes, 8 per octave (<= 12.5 % over-allocation): lattices of similar size - the points
    // of a concentration sweep - then reuse each other's blocks instead of reaching cudaMalloc,
    // which waits for the kernels of every other handle
    bytes = std::max<size_t>(bytes, 256);
    {
        size_t p2 = 256;
        while (p2 < bytes) p2 <<= 1;
        const size_t step = std::max<size_t>(p2 >> 4, 256);      // p2/2 .. p2 in 8 steps
        bytes = (bytes + step - 1) / step * step;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        int best = -1;
        for (int k = 0; k < (int)g_pool_parked.size(); ++k) {
            const PoolBlock& b = g_pool_parked[k];
            if (b.dev != dev || b.bytes != bytes) continue;
            if (best < 0 || b.bytes < g_pool_parked[best].bytes) best = k;
        }
        if (best >= 0) {
            PoolBlock b = g_pool_parked[best];
            g_pool_parked.erase(g_pool_parked.begin() + best);
            g_pool_parked_bytes -= b.bytes;
            g_pool_live[b.p] = b;
            *out = b.p;
            return cudaSuccess;
        }
        if (dev >= 0 && dev < STB_MAX_DEV) {
            arena_init(dev);
            if (void* q = arena_alloc(dev, bytes)) {
                g_pool_live[q] = PoolBlock{q, bytes, dev, false, true};
                *out = q;
                return cudaSuccess;
            }
        }
    }
    void* p = nullptr;
    static const bool verbose = getenv("STB_VERBOSE_POOL") != nullptr;
    timespec ts0, ts1;
    if (verbose) clock_gettime(CLOCK_MONOTONIC, &ts0);
    bool async = false;
    cudaError_t e = raw_malloc(&p, bytes, dev, &async);
    if (verbose) {
        clock_gettime(CLOCK_MONOTONIC, &ts1);
        fprintf(stderr, "[stb] pool miss: %s(%zu) took %.2f ms\n", async ? "cudaMallocAsync" : "cudaMalloc", bytes,
                (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);
    }
    if (e == cudaErrorMemoryAllocation) {
        // give parked memory back and retry once
        cudaGetLastError();
        std::vector<PoolBlock> drop;
        {
            std::lock_guard<std::mutex> lk(g_pool_mu);
            drop.swap(g_pool_parked);
            g_pool_parked_bytes = 0;
        }
        for (auto& b : drop) raw_free(b);
        e = raw_malloc(&p, bytes, dev, &async);
    }
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool_live[p] = PoolBlock{p, bytes, dev, async, false};
    *out = p;
    return cudaSuccess;
}
template <class T>
static cudaError_t pool_malloc(T** out, size_t bytes) { return pool_malloc((void**)out, bytes); }

// The caller guarantees that no kernel still uses the block (handles synchronise their stream
// before they release anything).
static void pool_free(void* p) {
    if (!p) return;
    PoolBlock b{nullptr, 0, 0, false, false};
    std::vector<PoolBlock> drop;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto it = g_pool_live.find(p);
        if (it == g_pool_live.end()) { b.p = p; }       // not ours: plain cudaFree below
        else {
            PoolBlock blk = it->second;
            g_pool_live.erase(it);
            if (blk.arena) {
                arena_free(blk.dev, blk.p, blk.bytes);
                return;
            }
            g_pool_parked.push_back(blk);
            g_pool_parked_bytes += blk.bytes;
            while (g_pool_parked_bytes > STB_POOL_CAP && !g_pool_parked.empty()) {
                drop.push_back(g_pool_parked.front());
                g_pool_parked_bytes -= g_pool_parked.front().bytes;
                g_pool_parked.erase(g_pool_parked.begin());
            }
        }
    }
    if (b.p) cudaFree(b.p);
    for (auto& d : drop) raw_free(d);
}

extern "C" int64_t stb_release_cached_memory(void) {
    std::vector<PoolBlock> drop;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        drop.swap(g_pool_parked);
        g_pool_parked_bytes = 0;
    }
    int64_t bytes = 0;
    for (auto& b : drop) { raw_free(b); bytes += (int64_t)b.bytes; }
    return bytes;
}

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            return STB_ECUDA;                                                         \
        }                                                                             \
    } while (0)

#define CKP(call)                                                                     \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            return nullptr;                                                           \
        }                                                                             \
    } while (0)

struct stb_system {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t n_ions = 0;
    int n[2] = {0, 0}, s[2] = {0, 0}, N = 0;
    int n_proc = 0;
    std::vector<stb_process> procs;
    std::vector<uint8_t> type;
    std::vector<int> local_idx, gs_ref, plane_of_ref, ref_of_plane;
    std::vector<int> gs_of_local[2];
    std::vector<int> local_of_rank_S;    // quad lattices: type-local index of the r-th sensitizer
    int* d_local_of_rank_S = nullptr;
    stb_process* d_procs = nullptr;      // device copy of the process table (strengths: see d_strengths)
    std::vector<int> len;  // [2][n]
    DevLattice hL;         // host copy (device pointers inside)
    DevLattice* dL = nullptr;
    // device buffers
    uint8_t* d_type = nullptr;
    int *d_local_idx = nullptr, *d_gs_ref = nullptr, *d_plane_of_ref = nullptr,
        *d_ref_of_plane = nullptr, *d_len = nullptr;
    int* d_gs_of_local[2] = {nullptr, nullptr};
    uint32_t* d_col[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    double* d_g[2][2][STB_MAXM] = {};
    int64_t* d_line_ptr[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int64_t pairs = 0, slots = 0, stream_bytes = 0;
    std::vector<int64_t> h_line_ptr[2][2];
    // batch
    int n_sys = 0;
    std::vector<double> h_blocks, h_strengths;
    double *d_blocks = nullptr, *d_strengths = nullptr;
    bool params_dirty = true;
    // scratch for rhs/jv
    double *d_y = nullptr, *d_v = nullptr, *d_out = nullptr, *d_F = nullptr, *d_tmp = nullptr;
    int64_t scratch_n = 0;
    float last_ms = 0.f, last_ms_pulse = 0.f, last_ms_relax = 0.f;
    std::vector<long long> last_prof;
    int last_G = 0, last_groups = 0;
    int64_t last_launches = 0;
    int sm_count = 148;
    int sm_budget = 148;     // SMs one solve of this handle may occupy (stb_set_option "sm_budget")
    bool built_with_cells = false;   // the cutoff lists came from the cell-list build
    std::vector<int> restart_list;         // systems of the last launch that asked for the 16-ion groups
    bool can16 = false, pcg_auto = true;   // 16-ion preconditioner groups possible / chosen by the restart rule
    bool avg_raw = false;            // the relaxation of a cumulative run: unclipped means (see k_cumtrapz)
    bool avg_cumulative = false;     // stb_integrate_dynamics returns the running time integral of the curves
    // integrator workspace: grow-only device buffers kept for the life of the handle, so a solve
    // issues no cudaMalloc / cudaFree (both synchronise the device and jitter by tens of ms)
    enum { WS_WORK, WS_Y0, WS_TMP, WS_T, WS_YOUT, WS_AVG, WS_NEXT, WS_STATS, WS_PROF, WS_PART, WS_BAR,
           WS_DYN_T, WS_DYN_TMP, WS_DYN_Y0, WS_DYN_YP, WS_DYN_YS, WS_DYN_AVG, WS_LIST, WS_COUNT };
    void* ws_ptr[WS_COUNT] = {};
    size_t ws_cap[WS_COUNT] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // cooperative (three-ion) process: reference-ordered triples on the host (export), ion-owned
    // lists on the device
    std::vector<uint32_t> h_coop_N;      // [Mc][3]
    std::vector<int32_t> h_coop_rows;    // [Mc][6]
    std::vector<double> h_coop_w;        // [Mc] strength * geometry
    void* d_coop[3][4] = {};             // per role: row_ptr, p1, p2, w
    int n_terms_pair[2] = {0, 0}, nf_pair[2] = {0, 0};
    // optimisation objective: experimental decay curves per state kind (host copies; uploaded lazily)
    std::vector<std::vector<double>> exp_t, exp_y;
    std::vector<double> exp_tail;
    bool exp_dirty = false;
    double *d_exp_t = nullptr, *d_exp_y = nullptr, *d_exp_tail = nullptr, *d_obj = nullptr;
    int64_t* d_exp_off = nullptr;
    int64_t exp_cap = 0, obj_cap = 0;
    // curves of the last stb_integrate_dynamics call (device): what stb_objective evaluates
    const double *last_avg = nullptr, *last_ts = nullptr;
    int last_dyn_n = 0;
    int64_t last_dyn_nt = 0;
};

// grow-only workspace slot of the handle
static cudaError_t ws_get(stb_system* S, int slot, size_t bytes, void** out) {
    bytes = std::max<size_t>(bytes, 16);
    if (bytes > S->ws_cap[slot]) {
        if (S->ws_ptr[slot]) pool_free(S->ws_ptr[slot]);
        S->ws_ptr[slot] = nullptr;
        S->ws_cap[slot] = 0;
        cudaError_t e = pool_malloc(&S->ws_ptr[slot], bytes);
        if (e != cudaSuccess) return e;
        S->ws_cap[slot] = bytes;
    }
    *out = S->ws_ptr[slot];
    return cudaSuccess;
}
static cudaError_t ws_events(stb_system* S) {
    if (S->ev0) return cudaSuccess;
    cudaError_t e = cudaEventCreate(&S->ev0);
    if (e != cudaSuccess) return e;
    return cudaEventCreate(&S->ev1);
}

extern "C" const char* stb_last_error(void) { return g_err.c_str(); }

extern "C" int stb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int stb_describe(char* buf, int64_t buf_len) {
    snprintf(buf, (size_t)buf_len,
             "simetuc_b200 sm_100a; integrator: persistent cooperative kernel, %d threads/CTA, "
             "GMRES(%d), BDF order<=%d; pair stream: dense f64 blocks (all-to-all) / sliced-ELL u32+f64 "
             "(cutoff)",
             STB_NT, STB_KRYLOV, STB_MAXORD);
    return STB_OK;
}

// Blocking copy ordered on the handle's own stream.  (cudaMemcpy uses the legacy default stream and
// waited behind the integrator kernels of other handles.)
static cudaError_t copy_on(cudaStream_t st, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

template <class T>
static cudaError_t upload(cudaStream_t st, T** dst, const std::vector<T>& v) {
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    cudaError_t e = pool_malloc((void**)dst, bytes);
    if (e != cudaSuccess) return e;
    if (!v.empty()) e = copy_on(st, *dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

static bool invert3(const double* m, double* inv) {
    double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                 m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (det == 0.0) return false;
    double id = 1.0 / det;
    inv[0] = (m[4] * m[8] - m[5] * m[7]) * id;
    inv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    inv[3] = (m[5] * m[6] - m[3] * m[8]) * id;
    inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    inv[6] = (m[3] * m[7] - m[4] * m[6]) * id;
    inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
    inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    return true;
}

extern "C" void stb_system_destroy(stb_system* S) {
    if (!S) return;
    cudaSetDevice(S->device);
    if (S->stream) cudaStreamSynchronize(S->stream);      // nothing may still use the pooled blocks
    pool_free(S->dL);
    pool_free(S->d_type);
    pool_free(S->d_local_idx);
    pool_free(S->d_local_of_rank_S);
    pool_free(S->d_procs);
    pool_free(S->d_gs_ref);
    pool_free(S->d_plane_of_ref);
    pool_free(S->d_ref_of_plane);
    pool_free(S->d_len);
    for (int T = 0; T < 2; ++T) {
        pool_free(S->d_gs_of_local[T]);
        for (int U = 0; U < 2; ++U) {
            pool_free(S->d_col[T][U]);
            pool_free(S->d_line_ptr[T][U]);
            for (int m = 0; m < STB_MAXM; ++m) pool_free(S->d_g[T][U][m]);
        }
    }
    pool_free(S->d_blocks);
    pool_free(S->d_strengths);
    pool_free(S->d_y);
    pool_free(S->d_v);
    pool_free(S->d_out);
    pool_free(S->d_F);
    pool_free(S->d_tmp);
    for (int k = 0; k < stb_system::WS_COUNT; ++k) pool_free(S->ws_ptr[k]);
    pool_free(S->d_exp_t);
    pool_free(S->d_exp_y);
    pool_free(S->d_exp_tail);
    pool_free(S->d_exp_off);
    pool_free(S->d_obj);
    for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 4; ++k) pool_free(S->d_coop[r][k]);
    if (S->ev0) cudaEventDestroy(S->ev0);
    if (S->ev1) cudaEventDestroy(S->ev1);
    if (S->stream) cudaStreamDestroy(S->stream);
    delete S;
}

// Field slots and rate terms of the process table (see stb_fields.cuh).
static int build_tables(stb_system* S, DevLattice& L) {
    // distinct multipolarities
    L.n_mult = 0;
    for (auto& p : S->procs) {
        bool seen = false;
        for (int m = 0; m < L.n_mult; ++m) seen |= L.mults[m] == p.mult;
        if (!seen) {
            if (L.n_mult == STB_MAXM) return fail(STB_EINVAL, "more than 3 distinct multipolarities");
            L.mults[L.n_mult++] = p.mult;
        }
    }
    if (L.n_mult == 0) { L.n_mult = 1; L.mults[0] = 6; }
    auto mult_index = [&](int mult) {
        for (int m = 0; m < L.n_mult; ++m) if (L.mults[m] == mult) return m;
        return 0;
    };
    for (int T = 0; T < 2; ++T) {
        for (int U = 0; U < 2; ++U) L.blk[T][U].n_slots = 0;
        L.terms[T].n_terms = 0;
    }
    auto slot_of = [&](int T, int U, int col, int mi) -> int {
        PairBlock& B = L.blk[T][U];
        for (int k = 0; k < B.n_slots; ++k)
            if (B.slot_col[k] == col && B.slot_mult[k] == mi) return k;
        if (B.n_slots == STB_MAXC) return -1;
        B.slot_col[B.n_slots] = col;
        B.slot_mult[B.n_slots] = mi;
        return B.n_slots++;
    };
    struct Raw { int T, U, slot, a, b, proc; };
    std::vector<Raw> raw;
    for (int p = 0; p < S->n_proc; ++p) {
        const stb_process& pr = S->procs[p];
        int dt = pr.donor_type, pt = pr.partner_type, mi = mult_index(pr.mult);
        if (S->s[dt] == 0 || S->s[pt] == 0) continue;
        // donor term on rows of type dt: field over partners pt, column fi
        int k1 = slot_of(dt, pt, pr.fi, mi);
        // acceptor term on rows of type pt: field over donors dt, column ii
        int k2 = slot_of(pt, dt, pr.ii, mi);
        if (k1 < 0 || k2 < 0) return fail(STB_EINVAL, "too many distinct field slots");
        raw.push_back({dt, pt, k1, pr.ii, pr.if_, p});
        raw.push_back({pt, dt, k2, pr.fi, pr.ff, p});
    }
    for (int T = 0; T < 2; ++T) {
        L.nf[T] = L.blk[T][0].n_slots + L.blk[T][1].n_slots;
    }
    for (auto& r : raw) {
        TermTable& tt = L.terms[r.T];
        if (tt.n_terms == STB_MAXT) return fail(STB_EINVAL, "too many ET terms");
        int k = tt.n_terms++;
        tt.fslot[k] = (short)(r.U == 0 ? r.slot : L.blk[r.T][0].n_slots + r.slot);
        tt.a[k] = (short)r.a;
        tt.b[k] = (short)r.b;
        tt.proc[k] = (short)r.proc;
    }
    L.foff[0] = 0;
    L.foff[1] = L.nf[0] * L.n[0];
    L.NF = L.nf[0] * L.n[0] + L.nf[1] * L.n[1];
    return STB_OK;
}

// Order of the sensitizers for the quad block preconditioner of the cluster integrator: strongly
// coupled ions (the closest pairs: energy migration goes with d^-6) are put next to each other,
// in quads of type-local indices 4m .. 4m+3.  Greedy: every sensitizer proposes its three nearest
// sensitizers, the proposed pairs are merged closest first while a group holds at most four ions,
// and the groups are packed into quads (4, 3+1, 2+2, 2+1+1, 1+1+1+1).  Only the ORDER of the ions
// inside the library changes; every result is reported in the reference's order.
// Returns false (identity order) for lattices it does not apply to.  out[r] = type-local index of
// the r-th sensitizer in lattice order.
// Cell grid for the cell-list pair build: cells no thinner than d_max along every axis (fractional
// coordinates of the periodic supercell, or the bounding box of an open lattice).  Returns false when
// the grid would be too coarse to pay off (a periodic axis needs >= 3 cells so that the 27
// neighbours are distinct).  cell_ions keeps the ions of a cell in ascending order.
static bool build_cell_grid(int n, const double* pos, const double* cell, const double* inv, double d_max,
                            int nc[3], std::vector<int>& cell_start, std::vector<int>& cell_ions,
                            std::vector<int>& ion_cell) {
    double lo[3] = {1e300, 1e300, 1e300}, ext[3] = {0, 0, 0};
    if (cell) {
        for (int c = 0; c < 3; ++c) {
            // height of the supercell along fractional axis c = 1 / |column c of inv|
            const double nrm = std::sqrt(inv[c] * inv[c] + inv[3 + c] * inv[3 + c] + inv[6 + c] * inv[6 + c]);
            nc[c] = (int)std::floor(1.0 / (nrm * d_max * (1.0 + 1e-9)));      // cells a hair wider than d_max
            if (nc[c] < 3) return false;
        }
    } else {
        double hi[3] = {-1e300, -1e300, -1e300};
        for (int q = 0; q < n; ++q)
            for (int c = 0; c < 3; ++c) {
                lo[c] = std::min(lo[c], pos[3 * q + c]);
                hi[c] = std::max(hi[c], pos[3 * q + c]);
            }
        for (int c = 0; c < 3; ++c) {
            ext[c] = std::max(hi[c] - lo[c], 1e-9);
            nc[c] = std::max(1, (int)std::floor(ext[c] / (d_max * (1.0 + 1e-9))));
        }
        if ((long long)nc[0] * nc[1] * nc[2] < 27) return false;
    }
    for (int c = 0; c < 3; ++c) nc[c] = std::min(nc[c], 256);
    const int n_cells = nc[0] * nc[1] * nc[2];
    ion_cell.resize(n);
    cell_start.assign(n_cells + 1, 0);
    for (int q = 0; q < n; ++q) {
        int idx[3];
        for (int c = 0; c < 3; ++c) {
            double f;
            if (cell) {
                f = pos[3 * q] * inv[c] + pos[3 * q + 1] * inv[3 + c] + pos[3 * q + 2] * inv[6 + c];
                f -= std::floor(f);
            } else {
                f = (pos[3 * q + c] - lo[c]) / ext[c];
            }
            idx[c] = std::min(nc[c] - 1, std::max(0, (int)(f * nc[c])));
        }
        ion_cell[q] = (idx[0] * nc[1] + idx[1]) * nc[2] + idx[2];
        cell_start[ion_cell[q] + 1]++;
    }
    for (int c = 0; c < n_cells; ++c) cell_start[c + 1] += cell_start[c];
    cell_ions.resize(n);
    std::vector<int> fill(cell_start.begin(), cell_start.end() - 1);
    for (int q = 0; q < n; ++q) cell_ions[fill[ion_cell[q]]++] = q;      // ascending inside a cell
    return true;
}

// Greedy packing of n items into groups of <= 4 strongly coupled items: the candidate pairs (a < b,
// d2 = squared distance; every item proposes its nearest items) are merged closest first while a group
// holds at most four items, and the groups are emitted as 4, 3+1, 2+2, 2+1+1, 1+1+1+1.
struct PairCand { double d2; int a, b; };
static std::vector<int> greedy_fours(int n, std::vector<PairCand>& cand) {
    std::sort(cand.begin(), cand.end(), [](const PairCand& x, const PairCand& y) {
        return x.d2 != y.d2 ? x.d2 < y.d2 : (x.a != y.a ? x.a < y.a : x.b < y.b);
    });
    std::vector<int> grp(n);
    std::vector<std::vector<int>> members(n);
    for (int a = 0; a < n; ++a) { grp[a] = a; members[a] = {a}; }
    for (const PairCand& c : cand) {
        const int ga = grp[c.a], gb = grp[c.b];
        if (ga == gb || members[ga].size() + members[gb].size() > 4) continue;
        for (int m : members[gb]) { grp[m] = ga; members[ga].push_back(m); }
        members[gb].clear();
    }
    std::vector<std::vector<int>> by_size[5];
    for (int a = 0; a < n; ++a)
        if (!members[a].empty()) by_size[members[a].size()].push_back(members[a]);
    std::vector<int> order;          // items in storage order
    order.reserve(n);
    auto take = [&](int sz) {
        std::vector<int> g = by_size[sz].back();
        by_size[sz].pop_back();
        order.insert(order.end(), g.begin(), g.end());
    };
    while (!by_size[4].empty()) take(4);
    while (!by_size[3].empty()) { take(3); if (!by_size[1].empty()) take(1); else break; }
    // (a triple without a single left over starts the tail: everything after it is packed as it comes)
    while (by_size[3].empty() && by_size[2].size() >= 2) { take(2); take(2); }
    while (by_size[3].empty() && !by_size[2].empty() && by_size[1].size() >= 2) { take(2); take(1); take(1); }
    for (int sz = 3; sz >= 1; --sz)
        while (!by_size[sz].empty()) take(sz);
    return order;
}

// Order of the sensitizers for the block preconditioners: strongly coupled ions (the closest pairs:
// energy migration goes with d^-6) are put next to each other, in quads of type-local indices
// 4m .. 4m+3, and neighbouring quads in groups of four, 16m .. 16m+15.  What counts is that no strong
// pair is cut by a group boundary (compact boxes of a k-d bisection cut too many: 3.97 Krylov iterations
// per solve against 2.02 on a 25 % Yb lattice, single-ion blocks 4.71), so both levels are greedy: every
// ion proposes its three nearest sensitizers, the pairs are merged closest first into quads; then every
// quad proposes its three nearest quads (distance = the closest pair of ions) and the quads are merged
// into groups the same way.  Neighbours: brute force up to 2048 sensitizers, through a uniform grid
// (positions required) above.  Only the ORDER of the ions inside the library changes; every result is
// reported in the reference's order.  Returns false (identity order) for lattices it does not apply to.
// out[r] = type-local index of the r-th sensitizer in lattice order.
static bool group_sensitizers(int n, const uint8_t* type, int sS, const double* dist, const double* pos,
                              const double* cell, std::vector<int>& out) {
    out.clear();
    if (const char* env = getenv("STB_QUADS")) if (atoi(env) == 0) return false;
    if (sS < 1 || 4 * sS > 12) return false;      // a quad block has at most 12 x 12 entries (Gauss-Jordan teams)
    std::vector<int> ions;
    for (int q = 0; q < n; ++q) if (type[q] == 0) ions.push_back(q);
    const int nS = (int)ions.size();
    if (nS < 2) return false;
    if (nS > 2048 && !pos) return false;
    double inv[9];
    const bool periodic = pos && cell && invert3(cell, inv);
    auto d2 = [&](int a, int b) -> double {
        if (dist) { const double d = dist[(size_t)ions[a] * n + ions[b]]; return d * d; }
        double v[3] = {pos[3 * ions[b]] - pos[3 * ions[a]], pos[3 * ions[b] + 1] - pos[3 * ions[a] + 1],
                       pos[3 * ions[b] + 2] - pos[3 * ions[a] + 2]};
        if (periodic) {      // nearest image by fractional rounding (good enough to find neighbours)
            double f[3];
            for (int c = 0; c < 3; ++c) f[c] = v[0] * inv[c] + v[1] * inv[3 + c] + v[2] * inv[6 + c];
            for (int c = 0; c < 3; ++c) f[c] -= std::nearbyint(f[c]);
            for (int c = 0; c < 3; ++c) v[c] = f[0] * cell[c] + f[1] * cell[3 + c] + f[2] * cell[6 + c];
        }
        return v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    };
    // ---- the K nearest sensitizers of every sensitizer ----------------------------------------------
    const int K = 8;
    std::vector<int> nbr((size_t)nS * K, -1);
    std::vector<double> nbd((size_t)nS * K, 1e300);
    auto offer = [&](int a, int b) {
        const double d = d2(a, b);
        if (!(d > 0.0)) return;
        double* bd = &nbd[(size_t)a * K];
        int* bn = &nbr[(size_t)a * K];
        if (!(d < bd[K - 1])) return;
        int k = K - 1;
        while (k > 0 && d < bd[k - 1]) { bd[k] = bd[k - 1]; bn[k] = bn[k - 1]; --k; }
        bd[k] = d; bn[k] = b;
    };
    if (nS <= 2048) {
        for (int a = 0; a < nS; ++a)
            for (int b = 0; b < nS; ++b) if (b != a) offer(a, b);
    } else {
        // uniform grid with ~4 sensitizers per cell, in fractional (periodic, wrapped) or bounding-box
        // coordinates; the 27 cells around an ion hold its nearest neighbours
        const int nc = std::max(3, (int)std::cbrt(nS / 4.0));
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        if (!periodic)
            for (int a = 0; a < nS; ++a)
                for (int c = 0; c < 3; ++c) {
                    lo[c] = std::min(lo[c], pos[3 * ions[a] + c]);
                    hi[c] = std::max(hi[c], pos[3 * ions[a] + c]);
                }
        std::vector<int> cidx((size_t)nS * 3), cell_of(nS), start((size_t)nc * nc * nc + 1, 0), items(nS);
        for (int a = 0; a < nS; ++a) {
            const double* p = pos + 3 * ions[a];
            for (int c = 0; c < 3; ++c) {
                double f;
                if (periodic) { f = p[0] * inv[c] + p[1] * inv[3 + c] + p[2] * inv[6 + c]; f -= std::floor(f); }
                else f = (p[c] - lo[c]) / std::max(hi[c] - lo[c], 1e-9);
                cidx[3 * a + c] = std::min(nc - 1, std::max(0, (int)(f * nc)));
            }
            cell_of[a] = (cidx[3 * a] * nc + cidx[3 * a + 1]) * nc + cidx[3 * a + 2];
            start[cell_of[a] + 1]++;
        }
        for (size_t c = 0; c + 1 < start.size(); ++c) start[c + 1] += start[c];
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int a = 0; a < nS; ++a) items[fill[cell_of[a]]++] = a;
        for (int a = 0; a < nS; ++a)
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dz = -1; dz <= 1; ++dz) {
                        int c3[3] = {cidx[3 * a] + dx, cidx[3 * a + 1] + dy, cidx[3 * a + 2] + dz};
                        bool in = true;
                        for (int c = 0; c < 3; ++c) {
                            if (periodic) c3[c] = (c3[c] + nc) % nc;
                            else if (c3[c] < 0 || c3[c] >= nc) in = false;
                        }
                        if (!in) continue;
                        const int cc = (c3[0] * nc + c3[1]) * nc + c3[2];
                        for (int i = start[cc]; i < start[cc + 1]; ++i) if (items[i] != a) offer(a, items[i]);
                    }
    }
    // ---- ions -> quads -----------------------------------------------------------------------------------
    std::vector<PairCand> cand;
    cand.reserve((size_t)nS * 3);
    for (int a = 0; a < nS; ++a)
        for (int k = 0; k < 3; ++k) {
            const int b = nbr[(size_t)a * K + k];
            if (b >= 0) cand.push_back({nbd[(size_t)a * K + k], std::min(a, b), std::max(a, b)});
        }
    const std::vector<int> ion_order = greedy_fours(nS, cand);
    // ---- quads -> groups of four quads: the last (possibly partial) quad stays last so that every other
    //      quad keeps its alignment -------------------------------------------------------------------------
    const int nq = (nS + 3) / 4;
    std::vector<int> quad_of(nS);
    for (int p = 0; p < nS; ++p) quad_of[ion_order[p]] = p / 4;
    std::vector<int> quad_order(nq);
    for (int q = 0; q < nq; ++q) quad_order[q] = q;
    bool group16 = nq > 1;
    if (const char* env = getenv("STB_GROUP16")) group16 = group16 && atoi(env) != 0;
    if (group16) {
        cand.clear();
        for (int q = 0; q + 1 < nq; ++q) {
            double best[3] = {1e300, 1e300, 1e300};
            int who[3] = {-1, -1, -1};
            for (int p = 4 * q; p < 4 * q + 4; ++p) {
                const int a = ion_order[p];
                for (int k = 0; k < K; ++k) {
                    const int b = nbr[(size_t)a * K + k];
                    if (b < 0) continue;
                    const int qb = quad_of[b];
                    if (qb == q || qb == nq - 1) continue;
                    const double d = nbd[(size_t)a * K + k];
                    int have = -1;
                    for (int m = 0; m < 3; ++m) if (who[m] == qb) have = m;
                    if (have >= 0) {
                        if (d < best[have]) best[have] = d;
                    } else if (d < best[2]) {
                        best[2] = d; who[2] = qb;
                    } else continue;
                    for (int m = 2; m > 0; --m)          // keep the three sorted
                        if (best[m] < best[m - 1]) { std::swap(best[m], best[m - 1]); std::swap(who[m], who[m - 1]); }
                    for (int m = 2; m > 0; --m)
                        if (best[m] < best[m - 1]) { std::swap(best[m], best[m - 1]); std::swap(who[m], who[m - 1]); }
                }
            }
            for (int m = 0; m < 3; ++m)
                if (who[m] >= 0) cand.push_back({best[m], std::min(q, who[m]), std::max(q, who[m])});
        }
        std::vector<int> go = greedy_fours(nq - 1, cand);
        for (int q = 0; q + 1 < nq; ++q) quad_order[q] = go[q];
    }
    std::vector<int> order;          // sensitizer ranks in storage order
    order.reserve(nS);
    for (int q : quad_order)
        for (int i = 4 * q; i < std::min(nS, 4 * q + 4); ++i) order.push_back(ion_order[i]);
    out.assign(nS, 0);
    for (int p = 0; p < nS; ++p) out[order[p]] = p;
    return true;
}

extern "C" stb_system* stb_system_create(int32_t device, int64_t n_ions, const uint8_t* ion_type,
                                         int32_t n_S_states, int32_t n_A_states,
                                         const double* dist_array, const double* positions,
                                         const double* cell, double d_max, int32_t n_proc,
                                         const stb_process* procs) {
    if (n_ions <= 0 || !ion_type) { fail(STB_EINVAL, "empty ion table"); return nullptr; }
    if ((dist_array == nullptr) == (positions == nullptr)) {
        fail(STB_EINVAL, "pass exactly one of dist_array / positions");
        return nullptr;
    }
    if (n_S_states < 0 || n_A_states < 0 || n_S_states > STB_MAX_STATES ||
        n_A_states > STB_MAX_STATES) {
        fail(STB_EINVAL, "states per ion must be in [0, 12]");
        return nullptr;
    }
    if (n_proc < 0 || n_proc > STB_MAX_PROC) { fail(STB_EINVAL, "too many processes"); return nullptr; }
    if (n_ions > 2000000) { fail(STB_EINVAL, "too many ions"); return nullptr; }
    CKP(cudaSetDevice(device));
    stb_system* S = new stb_system();
    // every failure below releases what the handle already owns (stream, pooled device blocks)
    double* d_geo = nullptr;
    auto die = [&]() -> stb_system* {
        if (d_geo) pool_free(d_geo);
        stb_system_destroy(S);
        return nullptr;
    };
#define CKS(call)                                                                     \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            return die();                                                             \
        }                                                                             \
    } while (0)

    S->device = device;
    S->n_ions = n_ions;
    S->s[0] = n_S_states;
    S->s[1] = n_A_states;
    S->n_proc = n_proc;
    S->procs.assign(procs, procs + n_proc);
    for (auto& p : S->procs) {
        int sd = S->s[p.donor_type & 1], sp = S->s[p.partner_type & 1];
        bool bad = p.donor_type < 0 || p.donor_type > 1 || p.partner_type < 0 || p.partner_type > 1 ||
                   p.ii < 0 || p.if_ < 0 || p.fi < 0 || p.ff < 0 || p.mult <= 0;
        // a process naming a state that does not exist is an error when it is active
        // (precalculate.py:325-345 raises LatticeError)
        if (!bad && stb_active_strength(p.strength) && sd > 0 && sp > 0)
            bad = p.ii >= sd || p.if_ >= sd || p.fi >= sp || p.ff >= sp;
        if (bad) {
            fail(STB_EINVAL, "The number of A or S states is lower than required by a process.");
            return die();
        }
    }
    cudaDeviceProp prop;
    CKS(cudaGetDeviceProperties(&prop, device));
    S->sm_count = prop.multiProcessorCount;
    CKS(cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking));

    // ---- ion table: type-local indices, ground states, layout maps (lattice.py:259-298) -------
    const int n = (int)n_ions;
    S->type.assign(ion_type, ion_type + n);
    S->local_idx.resize(n);
    S->gs_ref.resize(n);
    int cnt[2] = {0, 0}, gs = 0;
    for (int q = 0; q < n; ++q) {
        int T = S->type[q] != 0;
        S->type[q] = (uint8_t)T;
        S->local_idx[q] = cnt[T]++;
        S->gs_ref[q] = gs;
        gs += S->s[T];
    }
    S->n[0] = cnt[0];
    S->n[1] = cnt[1];
    S->N = gs;
    // quad lattices: strongly coupled sensitizers become neighbours in the type-local order
    bool quad = !(d_max > 0.0) && group_sensitizers(n, S->type.data(), S->s[0], dist_array, positions, cell,
                                                    S->local_of_rank_S);
    if (quad) {
        int r = 0;
        for (int q = 0; q < n; ++q)
            if (S->type[q] == 0) S->local_idx[q] = S->local_of_rank_S[r++];
    }
    for (int T = 0; T < 2; ++T) S->gs_of_local[T].assign(S->n[T], 0);
    for (int q = 0; q < n; ++q) S->gs_of_local[S->type[q]][S->local_idx[q]] = S->gs_ref[q];
    if (S->N == 0) { fail(STB_EINVAL, "No states: the number of energy states is zero."); return die(); }
    DevLattice& L = S->hL;
    memset(&L, 0, sizeof(L));
    L.n_ions = n;
    L.N = S->N;
    for (int T = 0; T < 2; ++T) { L.n[T] = S->n[T]; L.s[T] = S->s[T]; }
    L.off[0] = 0;
    L.off[1] = S->s[0] * S->n[0];
    L.n_proc = n_proc;
    L.quad = quad ? 1 : 0;
    // ions per preconditioner block of the cluster integrator: quads, or the 16-ion groups (blocks of <= 32
    // unknowns: 2-state sensitizers).  Measured (profiles/r02_groups.txt): on the reference's 240S/108A
    // benchmark lattice quads leave ~18 Krylov iterations per solve, groups 2.6: 507 -> 42 ms; on Yb-Tm
    // lattices (1.4-3.6 iterations with quads) groups save 15-22 % of the iterations and cost as much in
    // 32-pivot Gauss-Jordan sweeps, and whole groups per CTA unbalance small clusters (C2: 334 sensitizers
    // on 4 CTAs, 96 instead of 84 rows, pair rows no longer all resident: 722 -> 499 solves/s).  So a handle
    // starts with quads.  A single system that needs more than CL_RESTART_ITS iterations per solve over a
    // window of CL_RESTART_STEPS steps (within its first CL_RESTART_LAST steps) is stopped and solved again -
    // and the handle continues - with groups; a batch is never stopped, but switches for its next launches
    // when most of its systems needed more than CL_BATCH_ITS (integrate_cluster).  STB_PCG=4 / 16 fixes the choice.
    L.grp16 = quad && S->n[0] > 4;
    if (const char* env = getenv("STB_GROUP16")) if (atoi(env) == 0) L.grp16 = 0;
    S->can16 = L.grp16 && 16 * S->s[0] <= 32;
    L.pcg = 4;
    S->pcg_auto = true;
    if (const char* env = getenv("STB_PCG")) {
        L.pcg = (atoi(env) == 16 && S->can16) ? 16 : 4;
        S->pcg_auto = false;
    }
    S->plane_of_ref.resize(S->N);
    S->ref_of_plane.resize(S->N);
    for (int q = 0; q < n; ++q) {
        int T = S->type[q];
        for (int a = 0; a < S->s[T]; ++a) {
            int pl = L.off[T] + a * S->n[T] + S->local_idx[q];
            S->plane_of_ref[S->gs_ref[q] + a] = pl;
            S->ref_of_plane[pl] = S->gs_ref[q] + a;
        }
    }
    if (build_tables(S, L) != STB_OK) { return die(); }

    CKS(upload(S->stream, &S->d_type, S->type));
    CKS(upload(S->stream, &S->d_local_idx, S->local_idx));
    if (quad) CKS(upload(S->stream, &S->d_local_of_rank_S, S->local_of_rank_S));
    if (n_proc > 0) CKS(upload(S->stream, &S->d_procs, S->procs));
    CKS(upload(S->stream, &S->d_gs_ref, S->gs_ref));
    CKS(upload(S->stream, &S->d_plane_of_ref, S->plane_of_ref));
    CKS(upload(S->stream, &S->d_ref_of_plane, S->ref_of_plane));
    for (int T = 0; T < 2; ++T) CKS(upload(S->stream, &S->d_gs_of_local[T], S->gs_of_local[T]));
    L.plane_of_ref = S->d_plane_of_ref;
    L.ref_of_plane = S->d_ref_of_plane;

    // ---- geometry ------------------------------------------------------------------------------
    Geo G;
    memset(&G, 0, sizeof(G));
    G.n = n;
    G.d_max = d_max;
    if (dist_array) {
        CKS(pool_malloc((void**)&d_geo, sizeof(double) * (size_t)n * n));
        CKS(copy_on(S->stream, d_geo, dist_array, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
        G.dist = d_geo;
    } else {
        CKS(pool_malloc((void**)&d_geo, sizeof(double) * (size_t)n * 3));
        CKS(copy_on(S->stream, d_geo, positions, sizeof(double) * (size_t)n * 3, cudaMemcpyHostToDevice));
        G.pos = d_geo;
        if (cell) {
            G.periodic = 1;
            memcpy(G.cell, cell, sizeof(double) * 9);
            if (!invert3(G.cell, G.inv)) { fail(STB_EINVAL, "singular cell"); return die(); }
        }
    }
    // ---- cutoff lattices given by positions: cell lists (stb_build.cuh) ---------------------------
    CellGrid CG;
    memset(&CG, 0, sizeof(CG));
    bool use_cells = false;
    int *d_cell_start = nullptr, *d_cell_ions = nullptr, *d_ion_cell = nullptr;
    {
        std::vector<int> cell_start, cell_ions, ion_cell;
        int nc[3] = {1, 1, 1};
        bool want = d_max > 0.0 && positions != nullptr;
        if (const char* env = getenv("STB_CELLS")) want = want && atoi(env) != 0;
        if (want && build_cell_grid(n, positions, cell ? G.cell : nullptr, cell ? G.inv : nullptr, d_max, nc,
                                    cell_start, cell_ions, ion_cell)) {
            CKS(upload(S->stream, &d_cell_start, cell_start));
            CKS(upload(S->stream, &d_cell_ions, cell_ions));
            CKS(upload(S->stream, &d_ion_cell, ion_cell));
            CG.cell_start = d_cell_start;
            CG.cell_ions = d_cell_ions;
            CG.ion_cell = d_ion_cell;
            for (int c = 0; c < 3; ++c) CG.nc[c] = nc[c];
            CG.periodic = G.periodic;
            use_cells = true;
        }
    }
    auto free_cells = [&]() {
        pool_free(d_cell_start);
        pool_free(d_cell_ions);
        pool_free(d_ion_cell);
        d_cell_start = d_cell_ions = d_ion_cell = nullptr;
    };
    // ---- pass 1: row lengths ---------------------------------------------------------------------
    CKS(pool_malloc((void**)&S->d_len, sizeof(int) * 2 * (size_t)n));
    int blocks = std::min((n + 7) / 8, S->sm_count * 8);
    g_launches++;
    if (use_cells) k_count_pairs_cells<<<blocks, 256, 0, S->stream>>>(G, CG, S->d_type, S->d_len);
    else k_count_pairs<<<blocks, 256, 0, S->stream>>>(G, S->d_type, S->d_len);
    CKS(cudaGetLastError());
    S->len.resize(2 * (size_t)n);
    CKS(cudaMemcpyAsync(S->len.data(), S->d_len, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, S->stream));
    CKS(cudaStreamSynchronize(S->stream));
    // lanes per row from the mean row length (about 16 entries per lane)
    double mean_len = 0;
    for (int q = 0; q < n; ++q) mean_len += S->len[q] + S->len[n + q];
    mean_len /= n;
    int lpr = 1;
    while (lpr < 32 && lpr * 16 < mean_len) lpr *= 2;
    L.lpr = lpr;
    L.R = 32 / lpr;
    for (int T = 0; T < 2; ++T) L.n_slices[T] = (S->n[T] + L.R - 1) / L.R;
    // ---- all-to-all lattices: dense blocks, no partner index (stb_common.cuh) --------------------
    bool dense = !(d_max > 0.0);
    if (const char* env = getenv("STB_DENSE")) dense = dense && atoi(env) != 0;
    L.dense = dense ? 1 : 0;
    for (int U = 0; U < 2; ++U) L.ld[U] = (S->n[U] + 31) & ~31;
    if (dense) {
        FillDenseArgs FD;
        memset(&FD, 0, sizeof(FD));
        FD.G = G;
        FD.type = S->d_type;
        FD.local_idx = S->d_local_idx;
        FD.n_mult = L.n_mult;
        for (int m = 0; m < L.n_mult; ++m) FD.mults[m] = L.mults[m];
        S->pairs = 0;
        S->slots = 0;
        for (int T = 0; T < 2; ++T)
            for (int U = 0; U < 2; ++U) {
                FD.ld[U] = L.ld[U];
                L.blk[T][U].col = nullptr;
                L.blk[T][U].line_ptr = nullptr;
                if (L.blk[T][U].n_slots == 0 || S->n[T] == 0 || S->n[U] == 0) continue;
                const size_t cells = (size_t)S->n[T] * L.ld[U];
                for (int m = 0; m < L.n_mult; ++m) {
                    CKS(pool_malloc((void**)&S->d_g[T][U][m], cells * sizeof(double)));
                    CKS(cudaMemsetAsync(S->d_g[T][U][m], 0, cells * sizeof(double), S->stream));
                    L.blk[T][U].g[m] = S->d_g[T][U][m];
                    FD.g[T][U][m] = S->d_g[T][U][m];
                }
                S->slots += (int64_t)cells;
            }
        for (int q = 0; q < n; ++q)
            for (int U = 0; U < 2; ++U)
                if (L.blk[S->type[q]][U].n_slots > 0) S->pairs += S->len[(size_t)U * n + q];
        S->stream_bytes = S->slots * 8 * (int64_t)L.n_mult;
        g_launches++;
        k_fill_dense<<<blocks, 256, 0, S->stream>>>(FD);
        CKS(cudaGetLastError());
        CKS(cudaStreamSynchronize(S->stream));
    } else {
    // ---- slice line pointers ---------------------------------------------------------------------
    std::vector<int> ion_of_local[2];
    for (int T = 0; T < 2; ++T) ion_of_local[T].resize(S->n[T]);
    for (int q = 0; q < n; ++q) ion_of_local[S->type[q]][S->local_idx[q]] = q;
    S->pairs = 0;
    S->slots = 0;
    for (int T = 0; T < 2; ++T)
        for (int U = 0; U < 2; ++U) {
            std::vector<int64_t> lp(L.n_slices[T] + 1, 0);
            bool used = L.blk[T][U].n_slots > 0;
            for (int sl = 0; sl < L.n_slices[T]; ++sl) {
                int w = 0;
                for (int r = 0; r < L.R; ++r) {
                    int ql = sl * L.R + r;
                    if (ql >= S->n[T]) break;
                    int ln = S->len[(size_t)U * n + ion_of_local[T][ql]];
                    w = std::max(w, (ln + lpr - 1) / lpr);
                    if (used) S->pairs += ln;
                }
                lp[sl + 1] = lp[sl] + (used ? w : 0);
            }
            int64_t lines = lp.back();
            S->h_line_ptr[T][U] = lp;
            CKS(upload(S->stream, &S->d_line_ptr[T][U], lp));
            size_t slots = std::max<int64_t>(lines * 32, 1);
            CKS(pool_malloc((void**)&S->d_col[T][U], slots * sizeof(uint32_t)));
            CKS(cudaMemsetAsync(S->d_col[T][U], 0, slots * sizeof(uint32_t), S->stream));
            for (int m = 0; m < L.n_mult; ++m) {
                CKS(pool_malloc((void**)&S->d_g[T][U][m], slots * sizeof(double)));
                CKS(cudaMemsetAsync(S->d_g[T][U][m], 0, slots * sizeof(double), S->stream));
            }
            S->slots += lines * 32;
            L.blk[T][U].col = S->d_col[T][U];
            for (int m = 0; m < L.n_mult; ++m) L.blk[T][U].g[m] = S->d_g[T][U][m];
            L.blk[T][U].line_ptr = S->d_line_ptr[T][U];
        }
    S->stream_bytes = S->slots * (4 + 8 * (int64_t)L.n_mult);
    // ---- pass 2: fill ----------------------------------------------------------------------------
    FillArgs FA;
    memset(&FA, 0, sizeof(FA));
    FA.G = G;
    FA.type = S->d_type;
    FA.local_idx = S->d_local_idx;
    FA.lpr = L.lpr;
    FA.R = L.R;
    FA.n_mult = L.n_mult;
    for (int m = 0; m < L.n_mult; ++m) FA.mults[m] = L.mults[m];
    for (int T = 0; T < 2; ++T)
        for (int U = 0; U < 2; ++U) {
            FA.col[T][U] = S->d_col[T][U];
            FA.line_ptr[T][U] = S->d_line_ptr[T][U];
            for (int m = 0; m < L.n_mult; ++m) FA.g[T][U][m] = S->d_g[T][U][m];
        }
    // blocks whose fields are never used hold no lines: skip them in the fill through a
    // zero-width slice (pos / lpr stays inside the slice only when lines exist), so make the
    // kernel ignore them explicitly
    for (int T = 0; T < 2; ++T)
        for (int U = 0; U < 2; ++U)
            if (L.blk[T][U].n_slots == 0) FA.col[T][U] = nullptr;
    g_launches++;
    if (use_cells) k_fill_pairs_cells<<<std::min((n + 3) / 4, S->sm_count * 16), 128, 0, S->stream>>>(FA, CG);
    else k_fill_pairs<<<blocks, 256, 0, S->stream>>>(FA);
    CKS(cudaGetLastError());
    CKS(cudaStreamSynchronize(S->stream));
    }
    free_cells();
    S->built_with_cells = use_cells;
    pool_free(d_geo);
    d_geo = nullptr;
    L.blk_stride = 2 * (S->s[0] * S->s[0] + S->s[1] * S->s[1]);
    CKS(pool_malloc((void**)&S->dL, sizeof(DevLattice)));
    CKS(copy_on(S->stream, S->dL, &L, sizeof(DevLattice), cudaMemcpyHostToDevice));
    if (stb_set_batch(S, 1) != STB_OK) { stb_system_destroy(S); return nullptr; }
    return S;
}
#undef CKS

extern "C" int stb_system_info(const stb_system* S, int64_t out[8]) {
    if (!S) return fail(STB_EINVAL, "null system");
    out[0] = S->n_ions;
    out[1] = S->n[0];
    out[2] = S->n[1];
    out[3] = S->N;
    out[4] = S->pairs;
    out[5] = S->stream_bytes;
    out[6] = S->slots;
    out[7] = S->hL.lpr;
    return STB_OK;
}

extern "C" int stb_ground_states(const stb_system* S, int64_t* index_S_i, int64_t* index_A_j,
                                 uint64_t* initial_population) {
    if (!S) return fail(STB_EINVAL, "null system");
    if (initial_population) memset(initial_population, 0, sizeof(uint64_t) * S->N);
    for (int q = 0; q < S->n_ions; ++q) {
        int T = S->type[q];
        bool has = S->s[T] != 0;
        if (index_S_i) index_S_i[q] = (T == 0 && has) ? S->gs_ref[q] : -1;
        if (index_A_j) index_A_j[q] = (T == 1 && has) ? S->gs_ref[q] : -1;
        if (initial_population && has) initial_population[S->gs_ref[q]] = 1;
    }
    return STB_OK;
}

extern "C" int stb_set_batch(stb_system* S, int32_t n_sys) {
    if (!S || n_sys <= 0) return fail(STB_EINVAL, "n_sys must be positive");
    CK(cudaSetDevice(S->device));
    int per = S->hL.blk_stride;
    S->n_sys = n_sys;
    S->h_blocks.assign((size_t)n_sys * std::max(per, 1), 0.0);
    S->h_strengths.resize((size_t)n_sys * std::max(S->n_proc, 1));
    for (int b = 0; b < n_sys; ++b)
        for (int p = 0; p < S->n_proc; ++p) S->h_strengths[(size_t)b * S->n_proc + p] = S->procs[p].strength;
    pool_free(S->d_blocks);
    pool_free(S->d_strengths);
    S->d_blocks = S->d_strengths = nullptr;
    CK(pool_malloc((void**)&S->d_blocks, sizeof(double) * S->h_blocks.size()));
    CK(pool_malloc((void**)&S->d_strengths, sizeof(double) * S->h_strengths.size()));
    S->params_dirty = true;
    return STB_OK;
}

extern "C" int stb_set_blocks(stb_system* S, int32_t system, int32_t which, const double* blk_S,
                              const double* blk_A) {
    if (!S || which < 0 || which > 1 || system >= S->n_sys) return fail(STB_EINVAL, "bad arguments");
    int sS2 = S->s[0] * S->s[0], sA2 = S->s[1] * S->s[1], per = sS2 + sA2;
    for (int b = (system < 0 ? 0 : system); b < (system < 0 ? S->n_sys : system + 1); ++b) {
        double* dst = S->h_blocks.data() + (size_t)b * 2 * per + (size_t)which * per;
        if (blk_S && sS2) memcpy(dst, blk_S, sizeof(double) * sS2);
        if (blk_A && sA2) memcpy(dst + sS2, blk_A, sizeof(double) * sA2);
    }
    S->params_dirty = true;
    return STB_OK;
}

extern "C" int stb_set_strengths(stb_system* S, int32_t system, const double* strengths) {
    if (!S || !strengths || system >= S->n_sys) return fail(STB_EINVAL, "bad arguments");
    for (int p = 0; p < S->n_proc; ++p) {
        const stb_process& pr = S->procs[p];
        if (stb_active_strength(strengths[p]) && S->s[pr.donor_type] > 0 && S->s[pr.partner_type] > 0 &&
            (pr.ii >= S->s[pr.donor_type] || pr.if_ >= S->s[pr.donor_type] ||
             pr.fi >= S->s[pr.partner_type] || pr.ff >= S->s[pr.partner_type]))
            return fail(STB_EINVAL, "The number of A or S states is lower than required by a process.");
    }
    for (int b = (system < 0 ? 0 : system); b < (system < 0 ? S->n_sys : system + 1); ++b)
        memcpy(S->h_strengths.data() + (size_t)b * S->n_proc, strengths, sizeof(double) * S->n_proc);
    S->params_dirty = true;
    return STB_OK;
}

static int sync_params(stb_system* S) {
    if (!S->params_dirty) return STB_OK;
    CK(cudaMemcpyAsync(S->d_blocks, S->h_blocks.data(), sizeof(double) * S->h_blocks.size(),
                       cudaMemcpyHostToDevice, S->stream));
    CK(cudaMemcpyAsync(S->d_strengths, S->h_strengths.data(), sizeof(double) * S->h_strengths.size(),
                       cudaMemcpyHostToDevice, S->stream));
    CK(cudaStreamSynchronize(S->stream));
    S->params_dirty = false;
    return STB_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" int64_t stb_num_interactions(const stb_system* S, int32_t system) {
    if (!S || system < 0 || system >= S->n_sys) { fail(STB_EINVAL, "bad system"); return -1; }
    const int n = (int)S->n_ions;
    if (S->N <= 2) return 0;  // precalculate.py:304
    int np[2][2] = {{0, 0}, {0, 0}};
    for (int p = 0; p < S->n_proc; ++p)
        if (stb_active_strength(S->h_strengths[(size_t)system * S->n_proc + p]))
            np[S->procs[p].donor_type][S->procs[p].partner_type]++;
    int64_t M = 0;
    for (int q = 0; q < n; ++q) {
        int T = S->type[q];
        if (S->s[T] == 0) continue;
        for (int U = 0; U < 2; ++U)
            if (S->s[U]) M += (int64_t)np[T][U] * S->len[(size_t)U * n + q];
    }
    return M;
}

extern "C" int stb_export_interactions(const stb_system* S, int32_t system, uint32_t* N_indices,
                                       int32_t* et_rows, double* et_w, uint32_t* jac_indices) {
    int64_t M = stb_num_interactions(S, system);
    if (M < 0) return STB_EINVAL;
    if (M == 0) return STB_OK;
    CK(cudaSetDevice(S->device));
    const int n = (int)S->n_ions;
    std::vector<stb_process> pr = S->procs;
    int np[2][2] = {{0, 0}, {0, 0}};
    for (int p = 0; p < S->n_proc; ++p) {
        pr[p].strength = S->h_strengths[(size_t)system * S->n_proc + p];
        if (stb_active_strength(pr[p].strength)) np[pr[p].donor_type][pr[p].partner_type]++;
    }
    std::vector<int64_t> off(n);
    int64_t run = 0;
    for (int q = 0; q < n; ++q) {
        off[q] = run;
        int T = S->type[q];
        if (S->s[T] == 0) continue;
        for (int U = 0; U < 2; ++U)
            if (S->s[U]) run += (int64_t)np[T][U] * S->len[(size_t)U * n + q];
    }
    ExportArgs EA;
    memset(&EA, 0, sizeof(EA));
    int64_t* d_off = nullptr;
    stb_process* d_pr = nullptr;
    CK(upload(S->stream, &d_off, off));
    CK(upload(S->stream, &d_pr, pr));
    EA.L = S->dL;
    EA.type = S->d_type;
    EA.local_idx = S->d_local_idx;
    EA.gs_ref = S->d_gs_ref;
    EA.gs_of_local[0] = S->d_gs_of_local[0];
    EA.gs_of_local[1] = S->d_gs_of_local[1];
    EA.local_of_rank[0] = S->d_local_of_rank_S;      // null unless the sensitizers are stored in quads
    EA.local_of_rank[1] = nullptr;
    EA.len = S->d_len;
    EA.offset = d_off;
    EA.procs = d_pr;
    EA.n_proc = S->n_proc;
    EA.n = n;
    if (N_indices) CK(pool_malloc((void**)&EA.N_indices, sizeof(uint32_t) * 2 * M));
    if (et_rows) CK(pool_malloc((void**)&EA.et_rows, sizeof(int32_t) * 4 * M));
    if (et_w) CK(pool_malloc((void**)&EA.et_w, sizeof(double) * M));
    if (jac_indices) CK(pool_malloc((void**)&EA.jac, sizeof(uint32_t) * 6 * M));
    int blocks = std::min((n + 7) / 8, S->sm_count * 8);
    g_launches++;
    k_export<<<blocks, 256, 0, S->stream>>>(EA);
    CK(cudaGetLastError());
    if (N_indices) CK(cudaMemcpyAsync(N_indices, EA.N_indices, sizeof(uint32_t) * 2 * M, cudaMemcpyDeviceToHost, S->stream));
    if (et_rows) CK(cudaMemcpyAsync(et_rows, EA.et_rows, sizeof(int32_t) * 4 * M, cudaMemcpyDeviceToHost, S->stream));
    if (et_w) CK(cudaMemcpyAsync(et_w, EA.et_w, sizeof(double) * M, cudaMemcpyDeviceToHost, S->stream));
    if (jac_indices) CK(cudaMemcpyAsync(jac_indices, EA.jac, sizeof(uint32_t) * 6 * M, cudaMemcpyDeviceToHost, S->stream));
    CK(cudaStreamSynchronize(S->stream));
    pool_free(EA.N_indices);
    pool_free(EA.et_rows);
    pool_free(EA.et_w);
    pool_free(EA.jac);
    pool_free(d_off);
    pool_free(d_pr);
    return STB_OK;
}

// ---------------------------------------------------------------------------------------------
static int ensure_scratch(stb_system* S, int n) {
    if (S->scratch_n >= n) return STB_OK;
    pool_free(S->d_y); pool_free(S->d_v); pool_free(S->d_out); pool_free(S->d_F); pool_free(S->d_tmp);
    S->d_y = S->d_v = S->d_out = S->d_F = S->d_tmp = nullptr;
    size_t bytes = sizeof(double) * (size_t)n * S->N;
    CK(pool_malloc((void**)&S->d_y, bytes));
    CK(pool_malloc((void**)&S->d_v, bytes));
    CK(pool_malloc((void**)&S->d_out, bytes));
    CK(pool_malloc((void**)&S->d_tmp, bytes));
    CK(pool_malloc((void**)&S->d_F, sizeof(double) * (size_t)n * std::max(S->hL.NF, 1)));
    S->scratch_n = n;
    return STB_OK;
}

static dim3 op_grid(const stb_system* S, int n) {
    int total = S->hL.n_slices[0] + S->hL.n_slices[1];
    int bx = std::max(1, std::min((total + STB_NW - 1) / STB_NW, S->sm_count * 4));
    return dim3(bx, n, 1);
}

static int check_range(const stb_system* S, int first, int n) {
    if (!S) return fail(STB_EINVAL, "null system");
    if (first < 0 || n <= 0 || first + n > S->n_sys) return fail(STB_EINVAL, "system range outside the batch");
    return STB_OK;
}

extern "C" int stb_to_planes(const stb_system* S, int32_t n, const double* y_host, double* y_planes_dev) {
    if (!S || n <= 0) return fail(STB_EINVAL, "bad arguments");
    CK(cudaSetDevice(S->device));
    double* tmp = nullptr;
    size_t bytes = sizeof(double) * (size_t)n * S->N;
    CK(pool_malloc((void**)&tmp, bytes));
    CK(cudaMemcpyAsync(tmp, y_host, bytes, cudaMemcpyHostToDevice, S->stream));
    int64_t tot = (int64_t)n * S->N;
    g_launches++;
    k_ref_to_planes<<<(unsigned)((tot + 255) / 256), 256, 0, S->stream>>>(S->d_plane_of_ref, S->N, n, tmp, y_planes_dev);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(S->stream));
    pool_free(tmp);
    return STB_OK;
}

extern "C" int stb_from_planes(const stb_system* S, int32_t n, const double* y_planes_dev, double* y_host) {
    if (!S || n <= 0) return fail(STB_EINVAL, "bad arguments");
    CK(cudaSetDevice(S->device));
    double* tmp = nullptr;
    size_t bytes = sizeof(double) * (size_t)n * S->N;
    CK(pool_malloc((void**)&tmp, bytes));
    int64_t tot = (int64_t)n * S->N;
    g_launches++;
    k_planes_to_ref<<<(unsigned)((tot + 255) / 256), 256, 0, S->stream>>>(S->d_plane_of_ref, S->N, n, y_planes_dev, tmp);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(y_host, tmp, bytes, cudaMemcpyDeviceToHost, S->stream));
    CK(cudaStreamSynchronize(S->stream));
    pool_free(tmp);
    return STB_OK;
}

template <class K> static cudaError_t allow_max_smem(K kernel);

// The TMA-fed kernels (stb_tma.cuh) serve dense single-multipolarity blocks whose rows are read
// by whole warps; by default from 64 MB of pair data upwards (below that the stream is L2-resident
// and the register version has less fixed cost).  STB_TMA=0 / 1 forces the choice.
static bool use_tma(const stb_system* S) {
    const bool eligible = S->hL.dense && S->hL.n_mult == 1 && S->hL.lpr == 32;
    if (const char* env = getenv("STB_TMA")) return eligible && atoi(env) != 0;
    return eligible && S->stream_bytes >= (int64_t)64 << 20;
}

// one persistent CTA per SM; STB_TMA_GRID overrides the CTA count (tests: several row tiles per CTA)
static dim3 tma_grid(const stb_system* S, int n) {
    int bx = S->sm_count;
    if (const char* env = getenv("STB_TMA_GRID")) bx = std::max(1, atoi(env));
    return dim3(bx, n, 1);
}

// ring geometry of the TMA kernels: STB_TMA_KC columns per stage (multiple of 32), STB_TMA_STAGES
static size_t tma_setup(OpArgs& A) {
    int kc = STB_TMA_KC, st = STB_TMA_STAGES;
    if (const char* env = getenv("STB_TMA_KC")) kc = std::max(32, atoi(env) / 32 * 32);
    if (const char* env = getenv("STB_TMA_STAGES")) st = std::min(STB_TMA_MAX_STAGES, std::max(2, atoi(env)));
    while (st > 2 && tma_dyn_bytes(kc, st) > 220 * 1024) --st;
    A.tma_kc = kc; A.tma_stages = st;
    const char* diag = getenv("STB_TMA_DIAG");
    A.tma_diag = diag ? atoi(diag) : 0;
    return tma_dyn_bytes(kc, st) + 128;   // slack: see cl_choose
}

static int launch_rhs(stb_system* S, int first, int n, int pulse, const double* y, double* out, double* F) {
    OpArgs A = {};
    A.L = S->dL; A.blocks = S->d_blocks; A.strengths = S->d_strengths;
    A.first = first; A.pulse = pulse; A.y = y; A.v = nullptr; A.Fy = nullptr; A.out = out; A.F = F;
    g_launches++;
    if (use_tma(S)) {
        CK(allow_max_smem(k_op_tma<false>));
        const size_t dyn = tma_setup(A);
        k_op_tma<false><<<tma_grid(S, n), STB_TMA_NT, dyn, S->stream>>>(A);
    } else {
        k_rhs<<<op_grid(S, n), STB_NT, 0, S->stream>>>(A);
    }
    CK(cudaGetLastError());
    if (S->hL.n_coop) {
        g_launches++;
        k_coop_add<<<dim3((unsigned)std::min<int64_t>((S->n_ions + 7) / 8, S->sm_count * 8), n), 256, 0, S->stream>>>(
            S->dL, y, nullptr, out);
        CK(cudaGetLastError());
    }
    return STB_OK;
}

static int launch_jv(stb_system* S, int first, int n, int pulse, const double* y, const double* v,
                     const double* Fy, double* out) {
    OpArgs A = {};
    A.L = S->dL; A.blocks = S->d_blocks; A.strengths = S->d_strengths;
    A.first = first; A.pulse = pulse; A.y = y; A.v = v; A.Fy = Fy; A.out = out; A.F = nullptr;
    g_launches++;
    if (use_tma(S)) {
        CK(allow_max_smem(k_op_tma<true>));
        const size_t dyn = tma_setup(A);
        k_op_tma<true><<<tma_grid(S, n), STB_TMA_NT, dyn, S->stream>>>(A);
    } else {
        k_jv<<<op_grid(S, n), STB_NT, 0, S->stream>>>(A);
    }
    CK(cudaGetLastError());
    if (S->hL.n_coop) {
        g_launches++;
        k_coop_add<<<dim3((unsigned)std::min<int64_t>((S->n_ions + 7) / 8, S->sm_count * 8), n), 256, 0, S->stream>>>(
            S->dL, y, v, out);
        CK(cudaGetLastError());
    }
    return STB_OK;
}

extern "C" int stb_rhs(stb_system* S, int32_t first, int32_t n, int32_t pulse, const double* y, double* dy) {
    int rc = check_range(S, first, n);
    if (rc) return rc;
    CK(cudaSetDevice(S->device));
    if ((rc = sync_params(S))) return rc;
    if ((rc = ensure_scratch(S, n))) return rc;
    size_t bytes = sizeof(double) * (size_t)n * S->N;
    int64_t tot = (int64_t)n * S->N;
    unsigned gb = (unsigned)((tot + 255) / 256);
    CK(cudaMemcpyAsync(S->d_tmp, y, bytes, cudaMemcpyHostToDevice, S->stream));
    g_launches++;
    k_ref_to_planes<<<gb, 256, 0, S->stream>>>(S->d_plane_of_ref, S->N, n, S->d_tmp, S->d_y);
    if ((rc = launch_rhs(S, first, n, pulse, S->d_y, S->d_out, S->d_F))) return rc;
    g_launches++;
    k_planes_to_ref<<<gb, 256, 0, S->stream>>>(S->d_plane_of_ref, S->N, n, S->d_out, S->d_tmp);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dy, S->d_tmp, bytes, cudaMemcpyDeviceToHost, S->stream));
    CK(cudaStreamSynchronize(S->stream));
    return STB_OK;
}

extern "C" int stb_jv(stb_system* S, int32_t first, int32_t n, int32_t pulse, const double* y,
                      const double* v, double* jv) {
    int rc = check_range(S, first, n);
    if (rc) return rc;
    CK(cudaSetDevice(S->device));
    if ((rc = sync_params(S))) return rc;
    if ((rc = ensure_scratch(S, n))) return rc;
    size_t bytes = sizeof(double) * (size_t)n * S->N;
    int64_t tot = (int64_t)n * S->N;
    unsigned gb = (unsigned)((tot + 255) / 256);
    CK(cudaMemcpyAsync(S->d_tmp, y, bytes, cudaMemcpyHostToDevice, S->stream));
    g_launches++;
    k_ref_to_planes<<<gb, 256, 0, S->stream>>>(S->d_plane_of_ref, S->N, n, S->d_tmp, S->d_y);
    CK(cudaMemcpyAsync(S->d_tmp, v, bytes, cudaMemcpyHostToDevice, S->stream));
    g_launches++;
    k_ref_to_planes<<<gb, 256, 0, S->stream>>>(S->d_plane_of_ref, S->N, n, S->d_tmp, S->d_v);
    if ((rc = launch_rhs(S, first, n, pulse, S->d_y, S->d_out, S->d_F))) return rc;   // fields at y
    if ((rc = launch_jv(S, first, n, pulse, S->d_y, S->d_v, S->d_F, S->d_out))) return rc;
    g_launches++;
    k_planes_to_ref<<<gb, 256, 0, S->stream>>>(S->d_plane_of_ref, S->N, n, S->d_out, S->d_tmp);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(jv, S->d_tmp, bytes, cudaMemcpyDeviceToHost, S->stream));
    CK(cudaStreamSynchronize(S->stream));
    return STB_OK;
}

extern "C" int stb_rhs_dev(stb_system* S, int32_t first, int32_t n, int32_t pulse,
                           const double* y_planes, double* dy_planes, int32_t repeat, float* ms) {
    int rc = check_range(S, first, n);
    if (rc) return rc;
    CK(cudaSetDevice(S->device));
    if ((rc = sync_params(S))) return rc;
    if ((rc = ensure_scratch(S, n))) return rc;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    repeat = std::max(repeat, 1);
    CK(cudaEventRecord(e0, S->stream));
    for (int r = 0; r < repeat; ++r)
        if ((rc = launch_rhs(S, first, n, pulse, y_planes, dy_planes, S->d_F))) return rc;
    CK(cudaEventRecord(e1, S->stream));
    CK(cudaStreamSynchronize(S->stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, e0, e1));
    if (ms) *ms = t / repeat;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return STB_OK;
}

extern "C" int stb_jv_dev(stb_system* S, int32_t first, int32_t n, int32_t pulse,
                          const double* y_planes, const double* v_planes, double* jv_planes,
                          int32_t repeat, float* ms) {
    int rc = check_range(S, first, n);
    if (rc) return rc;
    CK(cudaSetDevice(S->device));
    if ((rc = sync_params(S))) return rc;
    if ((rc = ensure_scratch(S, n))) return rc;
    // fields at y once (what the integrator caches from its RHS pass)
    if ((rc = launch_rhs(S, first, n, pulse, y_planes, S->d_out, S->d_F))) return rc;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    repeat = std::max(repeat, 1);
    CK(cudaEventRecord(e0, S->stream));
    for (int r = 0; r < repeat; ++r)
        if ((rc = launch_jv(S, first, n, pulse, y_planes, v_planes, S->d_F, jv_planes))) return rc;
    CK(cudaEventRecord(e1, S->stream));
    CK(cudaStreamSynchronize(S->stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, e0, e1));
    if (ms) *ms = t / repeat;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return STB_OK;
}

// ---------------------------------------------------------------------------------------------
// Cluster integrator (small lattices): launch plan and driver.
// Opt both integrators into the full dynamic shared memory of an SM once per device: the attribute
// belongs to the function, not to a launch, and handles on several host threads launch concurrently.
template <class K>
static cudaError_t allow_max_smem(K kernel) {
    static std::mutex mu;
    static std::vector<std::pair<const void*, int>> done;      // (function, device) already set
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    for (auto& d : done)
        if (d.first == (const void*)kernel && d.second == dev) return cudaSuccess;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             232448 - (int)fa.sharedSizeBytes);
    if (e == cudaSuccess) done.emplace_back((const void*)kernel, dev);
    return e;
}

// STB_TIMING=1: host-side wall clock of the stages of an integrate call (development aid)
struct HostTimer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    std::string log;
    HostTimer() : on(getenv("STB_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* name) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        char buf[64];
        snprintf(buf, sizeof(buf), " %s=%.3f", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
        log += buf;
        t0 = t1;
    }
    ~HostTimer() { if (on) fprintf(stderr, "[stb timing]%s\n", log.c_str()); }
};

typedef void (*ClKernel)(ClArgs);
// the plain cluster integrator, or the instance with the cooperative field pass
// v2: the ion-major layout of the inner loops (stb_integrate_cl2.cuh)
static ClKernel cl_kernel(const stb_system* S, bool v2) {
    if (v2) return S->hL.n_coop ? (ClKernel)k_integrate_cl2_t<true> : (ClKernel)k_integrate_cl2_t<false>;
    return S->hL.n_coop ? (ClKernel)k_integrate_cl_t<true> : (ClKernel)k_integrate_cl_t<false>;
}

struct ClPlan {
    int G = 0, n_clusters = 0, resmask = 0, v_smem = 0, d_smem = 0;
    bool v2 = false;           // k_integrate_cl2_t (every CTA's ion states fit its threads)
    size_t dyn_bytes = 0;
    double resident_frac = 0.0;
    double est_us = 0.0;       // estimated time of one pair pass x rounds of the batch (see est_* below)
};

// Time model of one pair pass (microseconds), fitted to measurements on B200
// (profiles/r02_midsize.txt): the cluster integrator pays a small fixed latency per pass and reads
// the rows that do not fit in shared memory at ~20 GB/s per CTA; the grid integrator pays ~30 us of
// group barriers / reductions per pass and streams at ~27 GB/s per CTA (L2-resident data, G <= 64).
static double est_cluster_pass_us(int G, double stream_bytes, double resident_frac) {
    return 4.0 + 16.0 / G + stream_bytes * (1.0 - resident_frac) / G / 20e3;
}
static double est_grid_pass_us(int G, double stream_bytes) {
    return 28.0 + 0.1 * G + stream_bytes / (std::min(G, 64) * 27e3);
}

static cudaError_t cl_max_clusters(ClKernel kern, int G, size_t dyn_bytes, int sm_count, int* out) {
    cudaError_t e = allow_max_smem(kern);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(sm_count / G * G));
    cfg.blockDim = dim3(STB_NT);
    cfg.dynamicSmemBytes = dyn_bytes;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = (unsigned)G;
    at.val.clusterDim.y = 1;
    at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(out, kern, &cfg);
}

// Picks the cluster size and what lives in shared memory; false = use the cooperative-grid
// integrator (large lattices, cutoff lists, several multipolarities).
static bool cl_choose(stb_system* S, int n, ClPlan& plan) {
    const DevLattice& L = S->hL;
    if (!L.dense || L.n_mult != 1) return false;
    int forced = 0;
    if (const char* env = getenv("STB_GROUP_CTAS")) {
        forced = atoi(env);
        if (forced != 1 && forced != 2 && forced != 4 && forced != 8) return false;
    }
    // the ion-major variant (stb_integrate_cl2.cuh) is opt-in: measured slower than the element-major
    // kernel on the Yb-Tm sweeps (profiles/r02_integrator_notes.md)
    bool allow_v2 = false;
    if (const char* env = getenv("STB_CL_V")) allow_v2 = atoi(env) == 2;
    int want_res = 1;
    if (const char* env = getenv("STB_RESIDENT")) want_res = atoi(env) != 0;
    ClPlan best;
    double best_score = -1.0;
    for (int G = 8; G >= 1; G >>= 1) {
        if (forced && G != forced) continue;
        ClLayout Y;
        cl_layout(L, G, 0, 0, 0, Y);    // (ownership counts only)
        const bool v2 = allow_v2 && cl2_threads(L, Y.cntm[0], Y.cntm[1]) <= STB_NT;
        cudaFuncAttributes fa;
        ClKernel kern = cl_kernel(S, v2);
        if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) { cudaGetLastError(); continue; }
        const size_t cap = 232448 - fa.sharedSizeBytes - 256;  // 227 KB per CTA, minus the slack below
        cl_layout(L, G, 0, 0, 0, Y, v2);
        if ((Y.cntp[0] + Y.cntp[1]) / CL_RG > CL_MAXRG) continue;
        if ((size_t)Y.total * 8 > cap) continue;
        // shared memory goes first to the pair rows (largest block first), then to the Krylov
        // basis, then to the difference array
        int mask = 0, v_smem = 0, d_smem = 0;
        int64_t all = 0, res = 0;
        for (int T = 0; T < 2; ++T)
            for (int U = 0; U < 2; ++U) all += cl_block_doubles(L, Y, T, U);
        bool used[4] = {false, false, false, false};
        for (int round = 0; round < 4 && want_res; ++round) {
            int pick = -1, pick_sz = 0;
            for (int b = 0; b < 4; ++b) {
                int sz = cl_block_doubles(L, Y, b >> 1, b & 1);
                if (!used[b] && sz > pick_sz) { pick = b; pick_sz = sz; }
            }
            if (pick < 0) break;
            used[pick] = true;
            ClLayout Z;
            cl_layout(L, G, mask | (1 << pick), 0, 0, Z, v2);
            if ((size_t)Z.total * 8 <= cap) { mask |= 1 << pick; res += pick_sz; }
        }
        {
            // then the first 8 Krylov basis vectors (touched in every iteration), then the BDF
            // difference array (touched once per step), then the rest of the basis
            ClLayout Z;
            // the difference array is touched in every step, basis vectors beyond the first few
            // only by long linear solves: D first (with at least 4 vectors), then more vectors
            for (int c = 8; c >= 3 && !d_smem; --c) {
                cl_layout(L, G, mask, c, 1, Z, v2);
                if ((size_t)Z.total * 8 <= cap) { v_smem = c; d_smem = 1; }
            }
            for (int c = 8; c >= 1 && !v_smem; --c) {
                cl_layout(L, G, mask, c, 0, Z, v2);
                if ((size_t)Z.total * 8 <= cap) v_smem = c;
            }
            for (int c = CL_KRYLOV + 1; c > v_smem; --c) {
                cl_layout(L, G, mask, c, d_smem, Z, v2);
                if ((size_t)Z.total * 8 <= cap) { v_smem = c; break; }
            }
        }
        cl_layout(L, G, mask, v_smem, d_smem, Y, v2);
        // 128 B of slack: the compiler may align the dynamic shared memory base above the end of the
        // static variables (any `extern __shared__ __align__(n)` in the library does it for every
        // kernel), and the layout must still fit behind it
        const size_t bytes = (size_t)Y.total * sizeof(double) + 128;
        int maxc = 0;
        cudaError_t e = cl_max_clusters(kern, G, bytes, S->sm_count, &maxc);
        if (e != cudaSuccess || maxc < 1) { cudaGetLastError(); continue; }
        maxc = std::max(1, std::min(maxc, S->sm_budget / G));
        if (G > S->sm_budget && G > 1) continue;
        // rounds of the batch x time of one pass of one system
        const int ncl = std::min(maxc, n);
        const double rounds = std::ceil((double)n / ncl);
        const double rf = all > 0 ? (double)res / (double)all : 1.0;
        const double t_est = rounds * est_cluster_pass_us(G, (double)S->stream_bytes, rf);
        // between cluster sizes: rounds of the batch x time of one system (falls with G, less than
        // linearly; streaming the pair rows from L2 instead of shared memory roughly doubles a
        // pass) - fitted on batches, where every SM streams at once
        const double t_sys = (2.0 - rf) * (0.35 + 1.0 / G);
        const double score = 1.0 / (rounds * t_sys);
        if (score > best_score) {
            best_score = score;
            best.G = G; best.n_clusters = ncl; best.resmask = mask;
            best.v_smem = v_smem; best.d_smem = d_smem; best.dyn_bytes = bytes;
            best.resident_frac = rf;
            best.v2 = v2;
            best.est_us = t_est;
        }
    }
    if (best_score < 0.0) return false;
    plan = best;
    return true;
}

static int integrate_cluster(stb_system* S, const ClPlan& plan, int32_t first, int32_t n, int32_t pulse,
                             int64_t n_t, const double* t, const double* y0, double rtol, double atol,
                             int64_t nsteps, double* y_out, double* avg_out, stb_stats* stats, bool dev,
                             const std::vector<int>* list = nullptr) {
    // list: solve only these systems of the batch (local indices); the others keep the results of the
    // previous launch in the workspace (second pass of the quads -> groups rule)
    const int N = S->N;
    const DevLattice& L = S->hL;
    HostTimer tm;
    ClLayout Y;
    cl_layout(L, plan.G, plan.resmask, plan.v_smem, plan.d_smem, Y, plan.v2);
    const int64_t work_stride = (int64_t)(CL_KRYLOV + 1 + 8) * Y.PO;
    double *d_work = nullptr, *d_y0 = nullptr, *d_yout = nullptr, *d_t = nullptr, *d_tmp = nullptr,
           *d_avg = nullptr;
    int* d_next = nullptr;
    long long* d_prof = nullptr;
    stb_stats* d_stats = nullptr;
    const size_t out_bytes = sizeof(double) * (size_t)n * n_t * N;
    auto cleanup = [&]() {};   // every buffer is a workspace slot of the handle
#define CKC(call)                                                                     \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            cleanup();                                                                \
            return e_ == cudaErrorMemoryAllocation ? STB_ENOMEM : STB_ECUDA;          \
        }                                                                             \
    } while (0)
#define WSC(slot, ptr, bytes) CKC(ws_get(S, stb_system::slot, (bytes), (void**)&(ptr)))
    WSC(WS_WORK, d_work, sizeof(double) * (size_t)plan.n_clusters * plan.G * work_stride);
    const bool own_yout = !(dev && y_out);
    if (!dev) {
        WSC(WS_Y0, d_y0, sizeof(double) * (size_t)n * N);
        WSC(WS_TMP, d_tmp, sizeof(double) * (size_t)n * N);
        WSC(WS_T, d_t, sizeof(double) * n_t);
    }
    if (own_yout) WSC(WS_YOUT, d_yout, out_bytes);
    WSC(WS_NEXT, d_next, 2 * sizeof(int));        // work queue head, restart request
    WSC(WS_STATS, d_stats, sizeof(stb_stats) * n);
    WSC(WS_PROF, d_prof, sizeof(long long) * 64);
    CKC(ws_events(S));
    CKC(cudaMemsetAsync(d_prof, 0, sizeof(long long) * 64, S->stream));
    CKC(cudaMemsetAsync(d_next, 0, 2 * sizeof(int), S->stream));
    double* yout_p = own_yout ? d_yout : y_out;
    int* d_list = nullptr;
    if (!list) {
        CKC(cudaMemsetAsync(d_stats, 0, sizeof(stb_stats) * n, S->stream));
        CKC(cudaMemsetAsync(yout_p, 0, out_bytes, S->stream));
    } else {
        for (int k : *list) {
            CKC(cudaMemsetAsync(d_stats + k, 0, sizeof(stb_stats), S->stream));
            CKC(cudaMemsetAsync(yout_p + (size_t)k * n_t * N, 0, sizeof(double) * (size_t)n_t * N, S->stream));
        }
        WSC(WS_LIST, d_list, sizeof(int) * list->size());
        CKC(copy_on(S->stream, d_list, list->data(), sizeof(int) * list->size(), cudaMemcpyHostToDevice));
    }
    const double *t_p = t, *y0_p = y0;
    if (!dev) {
        CKC(cudaMemcpyAsync(d_t, t, sizeof(double) * n_t, cudaMemcpyHostToDevice, S->stream));
        CKC(cudaMemcpyAsync(d_tmp, y0, sizeof(double) * (size_t)n * N, cudaMemcpyHostToDevice, S->stream));
        int64_t tot = (int64_t)n * N;
        g_launches++;
        k_ref_to_planes<<<(unsigned)((tot + 255) / 256), 256, 0, S->stream>>>(S->d_plane_of_ref, N, n, d_tmp, d_y0);
        CKC(cudaGetLastError());
        t_p = d_t;
        y0_p = d_y0;
    }
    tm.mark("enqueue_inputs");
    ClArgs A;
    memset(&A, 0, sizeof(A));
    A.L = S->dL; A.blocks = S->d_blocks; A.strengths = S->d_strengths;
    A.first = first; A.n_sys = n; A.pulse = pulse; A.G = plan.G;
    A.resmask = plan.resmask; A.v_smem = plan.v_smem; A.d_smem = plan.d_smem;
    A.compact = plan.v2 ? 1 : 0;
    A.procs = S->d_procs;
    A.cst_base = 2;
    A.cst_slot = 3;
    A.piece_fix = CL_PIECE_FIX;
    A.lin_rms = CL_LIN_RMS;
    A.lin_tight_steps = CL_LIN_TIGHT_STEPS;
    if (const char* env = getenv("STB_LIN_TIGHT_STEPS")) A.lin_tight_steps = std::max(0, atoi(env));
    if (const char* env = getenv("STB_LIN_RMS")) A.lin_rms = std::max(1e-6, atof(env));
    if (const char* env = getenv("STB_PIECE_FIX")) A.piece_fix = std::max(0, atoi(env));
    if (const char* env = getenv("STB_CST_BASE")) A.cst_base = std::max(0, atoi(env));
    if (const char* env = getenv("STB_CST_SLOT")) A.cst_slot = std::max(1, atoi(env));
    A.pc_maxage = CL_PC_MAXAGE;
    A.pc_drift = CL_PC_DRIFT;
    if (const char* env = getenv("STB_PC_MAXAGE")) A.pc_maxage = std::max(1, atoi(env));
    if (const char* env = getenv("STB_PC_DRIFT")) A.pc_drift = std::max(0.0, atof(env));
    A.opts = CL_OPT_GUESS | CL_OPT_LAGPC | CL_OPT_QUAD;
    if (const char* env = getenv("STB_CL_OPTS")) A.opts = atoi(env);
    A.n_t = n_t; A.t = t_p; A.y0 = y0_p; A.rtol = rtol; A.atol = atol;
    A.nsteps = nsteps > 0 ? nsteps : 1000;
    A.y_out = yout_p; A.stats = d_stats; A.work = d_work; A.work_stride = work_stride;
    A.next_sys = d_next;
    // quads first; a lattice whose quads leave too many Krylov iterations is repeated with 16-ion groups
    A.sys_list = d_list;
    A.n_list = list ? (int)list->size() : 0;
    // (a single system: stopped and solved again as soon as it shows; a batch: never stopped - one stiff
    //  candidate of a parameter scan must not cost the others their launch - but if most of its systems
    //  needed more than CL_BATCH_ITS iterations per solve, the handle uses the groups from the next
    //  launch on)
    const bool may_switch = !list && L.quad && L.pcg == 4 && S->can16 && S->pcg_auto;
    const bool may_restart = may_switch && n == 1;
    A.restart = may_restart ? d_next + 1 : nullptr;
    A.restart_steps = CL_RESTART_STEPS;
    A.restart_its = CL_RESTART_ITS;
    A.prof = d_prof;
    S->last_G = plan.G; S->last_groups = plan.n_clusters;
    if (getenv("STB_VERBOSE"))
    {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, cl_kernel(S, plan.v2));
        ClLayout Z;
        cl_layout(L, plan.G, plan.resmask, 4, 1, Z, plan.v2);
        fprintf(stderr, "[stb] cluster integrator v%d: n=%d G=%d clusters=%d resmask=%x (%.0f%% of pair rows "
                "in smem) V_smem=%d D_smem=%d dyn_smem=%zu B own_pitch=%d static_smem=%zu B; with V=4 + D: %lld B "
                "of %zu B\n", plan.v2 ? 2 : 1, n, plan.G, plan.n_clusters,
                plan.resmask, 100.0 * plan.resident_frac, plan.v_smem, plan.d_smem, plan.dyn_bytes, Y.PO,
                (size_t)fa.sharedSizeBytes, (long long)Z.total * 8, (size_t)(232448 - fa.sharedSizeBytes - 256));
    }
    cudaEvent_t e0 = S->ev0, e1 = S->ev1;
    CKC(cudaEventRecord(e0, S->stream));
    {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(plan.n_clusters * plan.G));
        cfg.blockDim = dim3(STB_NT);
        cfg.dynamicSmemBytes = plan.dyn_bytes;
        cfg.stream = S->stream;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = (unsigned)plan.G;
        at.val.clusterDim.y = 1;
        at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        g_launches++;
        ClKernel kern = cl_kernel(S, plan.v2);
        CKC(allow_max_smem(kern));
        CKC(cudaLaunchKernelEx(&cfg, kern, A));
    }
    CKC(cudaEventRecord(e1, S->stream));
    tm.mark("launch");
    S->last_launches = dev ? 1 : 2;
    double* avg_p = avg_out;
    if (avg_out) {
        if (!dev) {
            WSC(WS_AVG, d_avg, sizeof(double) * (size_t)n * n_t * (L.s[0] + L.s[1]));
            avg_p = d_avg;
        }
        g_launches++;
        k_avg_populations<<<dim3((unsigned)n_t, n), 256, 0, S->stream>>>(S->dL, yout_p, n_t, avg_p,
                                                                             S->avg_raw ? 0 : 1);
        CKC(cudaGetLastError());
        S->last_launches += 1;
    }
    // Wait for the kernels HERE: a device-to-host copy into pageable memory issued behind a running
    // kernel blocks inside the driver with the context lock held, and the launches of every other
    // handle (other host threads) then wait for this solve to finish.
    CKC(cudaStreamSynchronize(S->stream));
    tm.mark("sync_kernels");
    if (may_restart) {
        int req = 0;
        CKC(copy_on(S->stream, &req, d_next + 1, sizeof(int), cudaMemcpyDeviceToHost));
        if (req) {
            std::vector<stb_stats> hs(n);
            CKC(copy_on(S->stream, hs.data(), d_stats, sizeof(stb_stats) * n, cudaMemcpyDeviceToHost));
            S->restart_list.clear();
            for (int k = 0; k < n; ++k) if (hs[k].status == 4) S->restart_list.push_back(k);
            CKC(cudaEventElapsedTime(&S->last_ms, e0, e1));
            if (!S->restart_list.empty()) return STB_RESTART_GROUPS;
        }
    } else if (may_switch && !pulse) {
        std::vector<stb_stats> hs(n);
        CKC(copy_on(S->stream, hs.data(), d_stats, sizeof(stb_stats) * n, cudaMemcpyDeviceToHost));
        // (the majority of the systems, not their mean: a few stiff candidates of a parameter scan do not make
        //  the groups pay for the others - C5: 184 -> 106 solves/s when its 8-candidate warm-up switched)
        int stiff = 0;
        for (const stb_stats& h : hs) stiff += (double)h.n_jv > CL_BATCH_ITS * (double)h.n_lin;
        if (2 * stiff > n) {
            S->hL.pcg = 16;
            ClPlan probe;
            if (!cl_choose(S, n, probe)) {       // the group blocks do not fit: quads it is, for good
                S->hL.pcg = 4;
                S->pcg_auto = false;
            }
            CKC(copy_on(S->stream, S->dL, &S->hL, sizeof(DevLattice), cudaMemcpyHostToDevice));
        }
    }
    if (avg_out) {
        if (!dev)
            CKC(cudaMemcpyAsync(avg_out, d_avg, sizeof(double) * (size_t)n * n_t * (L.s[0] + L.s[1]),
                                cudaMemcpyDeviceToHost, S->stream));
    }
    if (y_out && !dev) CKC(cudaMemcpyAsync(y_out, d_yout, out_bytes, cudaMemcpyDeviceToHost, S->stream));
    S->last_prof.assign((size_t)8 * n, 0);
    if (stats) CKC(cudaMemcpyAsync(stats, d_stats, sizeof(stb_stats) * n, cudaMemcpyDeviceToHost, S->stream));
    CKC(cudaStreamSynchronize(S->stream));
    tm.mark("d2h");
    CKC(cudaEventElapsedTime(&S->last_ms, e0, e1));
    if (tm.on) {
        char buf[48];
        snprintf(buf, sizeof(buf), " [kernel=%.3f]", S->last_ms);
        tm.log += buf;
    }
    if (getenv("STB_VERBOSE")) {
        long long pr[64];
        CKC(copy_on(S->stream, pr, d_prof, sizeof(pr), cudaMemcpyDeviceToHost));
        static const char* names[16] = {"stream", "stream_sync", "combine", "yk_barrier", "rhs_epilogue",
                                        "block_inv", "residual", "beta_reduce", "arnoldi_epi", "dots",
                                        "dots_reduce", "orth_push", "v_barrier", "newton_tail", "step_ctl",
                                        "loop_head"};
        // unit of each phase: p = per pair pass, n = per Newton iteration, k = per Krylov iteration,
        // s = per step
        static const char unit[16] = {'p', 'p', 'p', 'n', 'n', 'n', 'n', 'n', 'k', 'k', 'k', 'k', 'k', 'n', 's', 's'};
        double tot = 0;
        for (int k = 0; k < 16; ++k) tot += (double)pr[k];
        const double steps = std::max<double>(1, pr[17]), nrhs = std::max<double>(1, pr[18]), njv = std::max<double>(1, pr[19]);
        fprintf(stderr, "[stb] %.3f ms; cluster 0: %lld system(s), %.0f steps, %.0f rhs, %.0f jv, %.1f kcycles/step; "
                "cycles per unit (p = pass, n = Newton it, k = Krylov it, s = step):", S->last_ms, pr[16],
                steps, nrhs, njv, tot / steps / 1e3);
        for (int k = 0; k < 16; ++k) {
            const double den = unit[k] == 'p' ? nrhs + njv : unit[k] == 'n' ? nrhs : unit[k] == 'k' ? njv : steps;
            fprintf(stderr, " %s=%.0f/%c(%.0f%%)", names[k], pr[k] / den, unit[k], 100.0 * pr[k] / std::max(tot, 1.0));
        }
        if (plan.v2)
            fprintf(stderr, " | ion-major kernel, per Krylov it: products=%.0f butterfly=%.0f wait=%.0f orth_sums=%.0f", pr[20] / njv, pr[21] / njv, pr[22] / njv, pr[23] / njv);
        fprintf(stderr, "\n");
        if (!plan.v2) {
            fprintf(stderr, "[stb] pair pass per warp (cycles per pass | work-list steps):");
            for (int w = 0; w < STB_NW; ++w)
                fprintf(stderr, " %.0f|%lld", pr[24 + w] / (nrhs + njv), pr[24 + STB_NW + w + 1] - pr[24 + STB_NW + w]);
            fprintf(stderr, "\n");
        }
    }
    return STB_OK;
}

static int integrate_core(stb_system* S, int32_t first, int32_t n, int32_t pulse, int64_t n_t,
                          const double* t, const double* y0, double rtol, double atol,
                          int64_t nsteps, double* y_out, double* avg_out, stb_stats* stats,
                          bool dev) {
    int rc = check_range(S, first, n);
    if (rc) return rc;
    if (n_t < 1 || !t || !y0) return fail(STB_EINVAL, "bad time grid / initial state");
    if (!(rtol > 0.0) || !(atol >= 0.0)) return fail(STB_EINVAL, "rtol must be > 0 and atol >= 0");
    if (!dev)
        for (int64_t k = 1; k < n_t; ++k)
            if (!(t[k] > t[k - 1])) return fail(STB_EINVAL, "output times must increase strictly");
    CK(cudaSetDevice(S->device));
    if ((rc = sync_params(S))) return rc;
    const int N = S->N;
    const DevLattice& L = S->hL;
    const int NMinv = L.s[0] * L.s[0] * L.n[0] + L.s[1] * L.s[1] * L.n[1];

    // ---- small lattices: one thread-block cluster per system (stb_integrate_cl.cuh) ------------
    // ---- launch shape of the persistent cooperative grid: groups of G CTAs ------------------------
    const int max_ctas = std::max(1, std::min(S->sm_count, S->sm_budget));   // one CTA per SM
    const int total_slices = L.n_slices[0] + L.n_slices[1];
    // a CTA wants at least ~2 slices per warp; never more CTAs per group than that allows
    int G_work = std::max(1, total_slices / (2 * STB_NW));
    int G = std::max(1, std::min(max_ctas / std::min(n, max_ctas), G_work));
    // ---- small lattices: one thread-block cluster per system (stb_integrate_cl.cuh), unless the
    //      grid is estimated faster (pair rows that do not fit in shared memory: est_*_pass_us) ----
    {
        ClPlan plan;
        HostTimer tm;
        bool use_cl = cl_choose(S, n, plan);
        if (use_cl && (!S->hL.n_coop || L.lpr >= 32) && n <= plan.n_clusters) {
            // the whole batch runs at once: the latency of one solve decides
            const int groups = std::max(1, std::min(max_ctas / G, n));
            const double grid_us = std::ceil((double)n / groups) * est_grid_pass_us(G, (double)S->stream_bytes);
            if (grid_us < plan.est_us) use_cl = false;
        }
        if (const char* env = getenv("STB_CLUSTER")) {
            if (atoi(env) == 0 && (!S->hL.n_coop || L.lpr >= 32)) use_cl = false;
            else if (atoi(env) != 0 && plan.G > 0) use_cl = true;
        }
        tm.mark("cl_choose");
        if (use_cl) {
            rc = integrate_cluster(S, plan, first, n, pulse, n_t, t, y0, rtol, atol, nsteps, y_out,
                                   avg_out, stats, dev);
            if (rc != STB_RESTART_GROUPS) return rc;
            // the handle now uses the 16-ion groups (and keeps them): plan and run again
            // the systems that asked for it are solved again with the 16-ion groups.  All of them: the
            // handle keeps the groups; some of a batch: the next launch starts with quads again
            const float ms_quads = S->last_ms;          // the abandoned work counts in the reported time
            const std::vector<int> again = S->restart_list;
            const bool all = (int)again.size() == n;
            S->hL.pcg = 16;
            CK(copy_on(S->stream, S->dL, &S->hL, sizeof(DevLattice), cudaMemcpyHostToDevice));
            if (!cl_choose(S, (int)again.size(), plan)) {
                // the group blocks do not fit next to this lattice's rows: quads it is, for good
                S->hL.pcg = 4;
                S->pcg_auto = false;
                CK(copy_on(S->stream, S->dL, &S->hL, sizeof(DevLattice), cudaMemcpyHostToDevice));
                if (!cl_choose(S, n, plan)) return fail(STB_ESTATE, "no cluster plan");
                rc = integrate_cluster(S, plan, first, n, pulse, n_t, t, y0, rtol, atol, nsteps, y_out,
                                       avg_out, stats, dev);
                S->last_ms += ms_quads;
                return rc;
            }
            rc = integrate_cluster(S, plan, first, n, pulse, n_t, t, y0, rtol, atol, nsteps, y_out,
                                   avg_out, stats, dev, all ? nullptr : &again);
            if (!all) {
                S->hL.pcg = 4;
                CK(copy_on(S->stream, S->dL, &S->hL, sizeof(DevLattice), cudaMemcpyHostToDevice));
            }
            S->last_ms += ms_quads;
            return rc;
        }
    }
    if (S->hL.n_coop && L.lpr < 32)
        return fail(STB_EINVAL, "cooperative energy transfer in the grid integrator needs rows handled by whole "
                                "warps (lattices with >= 256 partners per ion); smaller lattices are integrated "
                                "by the cluster integrator (all-to-all pairs, one multipolarity)");
    // small systems: the gathered vector is staged in shared memory and barriers leave L1 warm;
    // if, with a few more CTAs per group, every CTA's own pair rows also fit in shared memory,
    // they are made resident there (the Krylov loop then runs out of shared memory)
    const size_t smem_cap = 200 * 1024;
    size_t stage_bytes = sizeof(double) * (size_t)N;
    int staged = stage_bytes <= 96 * 1024;
    if (const char* env = getenv("STB_STAGE")) staged = atoi(env) != 0 && stage_bytes <= 160 * 1024;
    int resident = 0;
    size_t dyn_bytes = staged ? stage_bytes + 128 : 0;     // slack: see cl_choose
    auto own_bytes = [&](int Gc) -> size_t {     // largest resident footprint of a CTA in a group of Gc
        size_t worst = 0;
        for (int c = 0; c < Gc; ++c) {
            size_t b = 0;
            for (int T = 0; T < 2; ++T) {
                int lo = (int)(((int64_t)L.n_slices[T] * c) / Gc), hi = (int)(((int64_t)L.n_slices[T] * (c + 1)) / Gc);
                for (int U = 0; U < 2; ++U) {
                    if (!L.blk[T][U].n_slots) continue;
                    const std::vector<int64_t>& lp = S->h_line_ptr[T][U];
                    b += (size_t)(lp[hi] - lp[lo]) * 32 * 12;
                }
                b += (size_t)(hi - lo + 1) * 8 * 2;
            }
            worst = std::max(worst, b);
        }
        return worst;
    };
    bool want_resident = staged && L.n_mult == 1 && !L.dense;
    if (const char* env = getenv("STB_RESIDENT")) want_resident = want_resident && atoi(env) != 0;
    if (const char* env = getenv("STB_GROUP_CTAS")) {
        G = std::max(1, std::min(atoi(env), max_ctas));
        if (want_resident && stage_bytes + own_bytes(G) + 64 <= smem_cap) resident = 1;
    } else if (want_resident) {
        // measured (profiles/r01_integrator_notes.md): widening the group just to become
        // resident costs more concurrency than it gains, so residency is used when it is free
        if (stage_bytes + own_bytes(G) + 64 <= smem_cap) resident = 1;
    }
    if (resident) dyn_bytes = stage_bytes + own_bytes(G) + 64 + 128;
    // dense rows read by whole warps: the pair stream goes through the TMA ring (stb_tma.cuh), the
    // gathered vector stays in global memory / L1.  STB_TMA_INT=0 keeps the register-fed passes.
    OpArgs tma_args = {};
    int tma = use_tma(S) ? 1 : 0;
    if (const char* env = getenv("STB_TMA_INT")) tma = tma && atoi(env) != 0;
    if (tma) {
        staged = 0;
        resident = 0;
        dyn_bytes = tma_setup(tma_args);
    }
    int n_groups = std::max(1, std::min(max_ctas / G, n));
    const int grid = n_groups * G;
    CK(allow_max_smem(k_integrate));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_integrate, STB_NT, dyn_bytes));
    if (occ < 1) return fail(STB_ECUDA, "integrator kernel does not fit on an SM");

    // group block preconditioner (stb_integrate.cuh): sensitizers stored in compact groups, dense S-S
    // block (the couplings inside a group are read from it), rows handled by >= 16 lanes
    int grp = 0;
    if (L.quad && L.grp16 && L.dense && L.n_mult == 1 && L.lpr >= 16 && !resident && L.s[0] >= 1 && STB_GDIM % L.s[0] == 0 &&
        (STB_GDIM / L.s[0]) % 4 == 0 && (STB_GDIM / L.s[0]) % L.R == 0 && S->d_procs && L.n_proc > 0)
        grp = STB_GDIM / L.s[0];
    if (const char* env = getenv("STB_GRID_GROUPS")) if (atoi(env) == 0) grp = 0;
    const int NGinv = grp ? ((L.n[0] + grp - 1) / grp + G + 1) * STB_GDIM * STB_GDIM : 0;     // group slots: see ch_slot
    const int64_t work_stride = stb_work_doubles(N, L.NF, NMinv + NGinv);
    double *d_work = nullptr, *d_y0 = nullptr, *d_yout = nullptr, *d_t = nullptr, *d_part = nullptr,
           *d_tmp = nullptr, *d_avg = nullptr;
    unsigned long long* d_bar = nullptr;
    stb_stats* d_stats = nullptr;
    long long* d_prof = nullptr;
    size_t out_bytes = sizeof(double) * (size_t)n * n_t * N;
    auto cleanup = [&]() {};   // every buffer is a workspace slot of the handle
#define CKI(call)                                                                     \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            cleanup();                                                                \
            return e_ == cudaErrorMemoryAllocation ? STB_ENOMEM : STB_ECUDA;          \
        }                                                                             \
    } while (0)
#define WSI(slot, ptr, bytes) CKI(ws_get(S, stb_system::slot, (bytes), (void**)&(ptr)))
    WSI(WS_WORK, d_work, sizeof(double) * (size_t)n_groups * work_stride);
    const bool own_yout = !(dev && y_out);
    if (!dev) {
        WSI(WS_Y0, d_y0, sizeof(double) * (size_t)n * N);
        WSI(WS_TMP, d_tmp, sizeof(double) * (size_t)n * N);
        WSI(WS_T, d_t, sizeof(double) * n_t);
    }
    if (own_yout) WSI(WS_YOUT, d_yout, out_bytes);
    WSI(WS_PART, d_part, sizeof(double) * (size_t)n_groups * 2 * G * STB_KRED);
    WSI(WS_BAR, d_bar, sizeof(unsigned long long) * ((size_t)n_groups + (size_t)n_groups * G));    // + neighbour flags
    WSI(WS_STATS, d_stats, sizeof(stb_stats) * n);
    WSI(WS_PROF, d_prof, sizeof(long long) * 8 * n);
    CKI(ws_events(S));
    CKI(cudaMemsetAsync(d_prof, 0, sizeof(long long) * 8 * n, S->stream));
    CKI(cudaMemsetAsync(d_bar, 0, sizeof(unsigned long long) * ((size_t)n_groups + (size_t)n_groups * G), S->stream));
    CKI(cudaMemsetAsync(d_stats, 0, sizeof(stb_stats) * n, S->stream));
    CKI(cudaMemsetAsync(d_work, 0, sizeof(double) * (size_t)n_groups * work_stride, S->stream));
    double* yout_p = own_yout ? d_yout : y_out;
    CKI(cudaMemsetAsync(yout_p, 0, out_bytes, S->stream));
    const double *t_p = t, *y0_p = y0;
    if (!dev) {
        CKI(cudaMemcpyAsync(d_t, t, sizeof(double) * n_t, cudaMemcpyHostToDevice, S->stream));
        CKI(cudaMemcpyAsync(d_tmp, y0, sizeof(double) * (size_t)n * N, cudaMemcpyHostToDevice, S->stream));
        int64_t tot = (int64_t)n * N;
        g_launches++;
        k_ref_to_planes<<<(unsigned)((tot + 255) / 256), 256, 0, S->stream>>>(S->d_plane_of_ref, N, n, d_tmp, d_y0);
        CKI(cudaGetLastError());
        t_p = d_t;
        y0_p = d_y0;
    }

    IntArgs A;
    A.L = S->dL; A.blocks = S->d_blocks; A.strengths = S->d_strengths;
    A.first = first; A.n_sys = n; A.pulse = pulse; A.G = G; A.n_groups = n_groups;
    A.n_t = n_t; A.t = t_p; A.y0 = y0_p; A.rtol = rtol; A.atol = atol;
    A.nsteps = nsteps > 0 ? nsteps : 1000;
    A.y_out = yout_p; A.stats = d_stats; A.work = d_work; A.work_stride = work_stride;
    A.bar = d_bar; A.partials = d_part; A.NMinv = NMinv; A.prof = d_prof; A.staged = staged; A.resident = resident;
    A.tma = tma; A.tma_kc = tma ? tma_args.tma_kc : 0; A.tma_stages = tma ? tma_args.tma_stages : 0;
    A.guess = 1;
    A.grp = grp; A.NGinv = NGinv; A.procs = S->d_procs;
    A.lag = 1;
    A.a_ctas = 1;
    if (const char* env = getenv("STB_A_CTAS")) A.a_ctas = atoi(env) != 0;
    if (const char* env = getenv("STB_GRID_LAG")) A.lag = atoi(env) != 0;
    A.lin_rms = GRID_LIN_RMS;
    A.lin_tight_steps = CL_LIN_TIGHT_STEPS;
    if (const char* env = getenv("STB_LIN_RMS_GRID")) A.lin_rms = std::max(1e-6, atof(env));
    if (const char* env = getenv("STB_LIN_TIGHT_STEPS")) A.lin_tight_steps = std::max(0, atoi(env));
    if (const char* env = getenv("STB_CL_OPTS")) A.guess = atoi(env) & CL_OPT_GUESS;
    S->last_G = G; S->last_groups = n_groups;
    cudaEvent_t e0 = S->ev0, e1 = S->ev1;
    CKI(cudaEventRecord(e0, S->stream));
    void* kargs[] = {&A};
    g_launches++;
    CKI(cudaLaunchCooperativeKernel((void*)k_integrate, dim3(grid), dim3(STB_NT), kargs, dyn_bytes, S->stream));
    CKI(cudaEventRecord(e1, S->stream));
    S->last_launches = dev ? 1 : 2;
    double* avg_p = avg_out;
    if (avg_out) {
        if (!dev) {
            WSI(WS_AVG, d_avg, sizeof(double) * (size_t)n * n_t * (L.s[0] + L.s[1]));
            avg_p = d_avg;
        }
        g_launches++;
        k_avg_populations<<<dim3((unsigned)n_t, n), 256, 0, S->stream>>>(S->dL, yout_p, n_t, avg_p,
                                                                             S->avg_raw ? 0 : 1);
        CKI(cudaGetLastError());
        S->last_launches += 1;
    }
    CKI(cudaStreamSynchronize(S->stream));     // before any pageable device-to-host copy, see integrate_cluster
    if (avg_out) {
        if (!dev)
            CKI(cudaMemcpyAsync(avg_out, d_avg, sizeof(double) * (size_t)n * n_t * (L.s[0] + L.s[1]),
                                cudaMemcpyDeviceToHost, S->stream));
    }
    if (y_out && !dev) CKI(cudaMemcpyAsync(y_out, d_yout, out_bytes, cudaMemcpyDeviceToHost, S->stream));
    S->last_prof.assign((size_t)8 * n, 0);
    CKI(cudaMemcpyAsync(S->last_prof.data(), d_prof, sizeof(long long) * 8 * n, cudaMemcpyDeviceToHost, S->stream));
    if (stats) CKI(cudaMemcpyAsync(stats, d_stats, sizeof(stb_stats) * n, cudaMemcpyDeviceToHost, S->stream));
    CKI(cudaStreamSynchronize(S->stream));
    CKI(cudaEventElapsedTime(&S->last_ms, e0, e1));
    return STB_OK;
}

extern "C" int stb_integrate(stb_system* S, int32_t first, int32_t n, int32_t pulse, int64_t n_t,
                             const double* t, const double* y0, double rtol, double atol,
                             int64_t nsteps, double* y_out, double* avg_out, stb_stats* stats) {
    return integrate_core(S, first, n, pulse, n_t, t, y0, rtol, atol, nsteps, y_out, avg_out, stats, false);
}

extern "C" int stb_integrate_dev(stb_system* S, int32_t first, int32_t n, int32_t pulse, int64_t n_t,
                                 const double* t_dev, const double* y0_planes_dev, double rtol,
                                 double atol, int64_t nsteps, double* y_out_dev, double* avg_out_dev,
                                 stb_stats* stats) {
    return integrate_core(S, first, n, pulse, n_t, t_dev, y0_planes_dev, rtol, atol, nsteps,
                          y_out_dev, avg_out_dev, stats, true);
}

// pulse + relaxation; dev: every array pointer except the statistics is a device pointer
// (t_pulse, t_sol; y0 in the plane layout; outputs in the reference layout) and nothing is copied
static int integrate_dynamics_core(stb_system* S, int32_t first, int32_t n, int64_t n_tp,
                                   const double* t_pulse, int64_t n_t, const double* t_sol,
                                   const double* y0, double rtol, double atol, int64_t nsteps_pulse,
                                   int64_t nsteps, double* y_pulse_out, double* y_out, double* avg_out,
                                   stb_stats* stats_pulse, stb_stats* stats_relax, bool dev) {
    int rc = check_range(S, first, n);
    if (rc) return rc;
    if (n_tp < 2 || n_t < 1 || !t_pulse || !t_sol || !y0)
        return fail(STB_EINVAL, "bad time grids / initial state");
    if (!dev) {
        for (int64_t k = 1; k < n_tp; ++k)
            if (!(t_pulse[k] > t_pulse[k - 1])) return fail(STB_EINVAL, "output times must increase strictly");
        for (int64_t k = 1; k < n_t; ++k)
            if (!(t_sol[k] > t_sol[k - 1])) return fail(STB_EINVAL, "output times must increase strictly");
    }
    CK(cudaSetDevice(S->device));
    HostTimer tm;
    const int N = S->N;
    const int kinds = S->hL.s[0] + S->hL.s[1];
    double *d_t = nullptr, *d_tmp = nullptr, *d_y0 = nullptr, *d_yp = nullptr, *d_ys = nullptr, *d_avg = nullptr;
    const size_t vec_bytes = sizeof(double) * (size_t)n * N;
    const int64_t tot = (int64_t)n * N;
    const double *tp_p = t_pulse, *ts_p = t_sol;
    CK(ws_get(S, stb_system::WS_DYN_Y0, vec_bytes, (void**)&d_y0));
    if (dev) {
        // the relaxation overwrites its start vector's buffer: keep the caller's y0 intact
        d_yp = y_pulse_out;
        d_ys = y_out;
        d_avg = avg_out;
    } else {
        CK(ws_get(S, stb_system::WS_DYN_T, sizeof(double) * (size_t)(n_tp + n_t), (void**)&d_t));
        CK(ws_get(S, stb_system::WS_DYN_TMP, vec_bytes, (void**)&d_tmp));
        if (y_out) CK(ws_get(S, stb_system::WS_DYN_YS, vec_bytes * n_t, (void**)&d_ys));
        if (avg_out || !S->exp_t.empty())
            CK(ws_get(S, stb_system::WS_DYN_AVG, sizeof(double) * (size_t)n * n_t * kinds, (void**)&d_avg));
        CK(cudaMemcpyAsync(d_t, t_pulse, sizeof(double) * n_tp, cudaMemcpyHostToDevice, S->stream));
        CK(cudaMemcpyAsync(d_t + n_tp, t_sol, sizeof(double) * n_t, cudaMemcpyHostToDevice, S->stream));
        CK(cudaMemcpyAsync(d_tmp, y0, vec_bytes, cudaMemcpyHostToDevice, S->stream));
        g_launches++;
        k_ref_to_planes<<<(unsigned)((tot + 255) / 256), 256, 0, S->stream>>>(S->d_plane_of_ref, N, n, d_tmp, d_y0);
        CK(cudaGetLastError());
        tp_p = d_t;
        ts_p = d_t + n_tp;
    }
    if (!d_yp) CK(ws_get(S, stb_system::WS_DYN_YP, vec_bytes * n_tp, (void**)&d_yp));
    tm.mark("inputs");
    rc = integrate_core(S, first, n, 1, n_tp, tp_p, dev ? y0 : d_y0, rtol, atol, nsteps_pulse, d_yp, nullptr,
                        stats_pulse, true);
    if (rc) return rc;
    const float ms_pulse = S->last_ms;
    const int64_t launches_pulse = S->last_launches;
    tm.mark("pulse");
    // the state at the end of the pulse starts the relaxation (simulations.py:1148-1157)
    g_launches++;
    k_row_to_planes<<<(unsigned)((tot + 255) / 256), 256, 0, S->stream>>>(
        S->d_plane_of_ref, N, n, d_yp + (n_tp - 1) * (int64_t)N, n_tp * (int64_t)N, d_y0);
    CK(cudaGetLastError());
    S->avg_raw = S->avg_cumulative && d_avg;     // k_cumtrapz clips after integrating
    rc = integrate_core(S, first, n, 0, n_t, ts_p, d_y0, rtol, atol, nsteps, d_ys, d_avg, stats_relax, true);
    S->avg_raw = false;
    if (rc) return rc;
    if (S->avg_cumulative && d_avg) {
        if (!S->exp_t.empty())
            return fail(STB_ESTATE, "avg_cumulative and the device objective (stb_set_experiment) exclude each other");
        g_launches++;
        k_cumtrapz<<<(unsigned)((n * kinds + 127) / 128), 128, 0, S->stream>>>(d_avg, ts_p, n_t, kinds, n);
        CK(cudaGetLastError());
    }
    tm.mark("relax");
    S->last_avg = d_avg;
    S->last_ts = ts_p;
    S->last_dyn_n = n;
    S->last_dyn_nt = n_t;
    S->last_ms_pulse = ms_pulse;
    S->last_ms_relax = S->last_ms;
    S->last_ms += ms_pulse;
    S->last_launches += launches_pulse + (dev ? 1 : 2);
    if (!dev) {
        if (y_pulse_out) CK(cudaMemcpyAsync(y_pulse_out, d_yp, vec_bytes * n_tp, cudaMemcpyDeviceToHost, S->stream));
        if (y_out) CK(cudaMemcpyAsync(y_out, d_ys, vec_bytes * n_t, cudaMemcpyDeviceToHost, S->stream));
        if (avg_out)
            CK(cudaMemcpyAsync(avg_out, d_avg, sizeof(double) * (size_t)n * n_t * kinds, cudaMemcpyDeviceToHost, S->stream));
        CK(cudaStreamSynchronize(S->stream));
        tm.mark("d2h");
    }
    return STB_OK;
}

extern "C" int stb_integrate_dynamics(stb_system* S, int32_t first, int32_t n, int64_t n_tp,
                                      const double* t_pulse, int64_t n_t, const double* t_sol,
                                      const double* y0, double rtol, double atol, int64_t nsteps_pulse,
                                      int64_t nsteps, double* y_pulse_out, double* y_out, double* avg_out,
                                      stb_stats* stats_pulse, stb_stats* stats_relax) {
    return integrate_dynamics_core(S, first, n, n_tp, t_pulse, n_t, t_sol, y0, rtol, atol, nsteps_pulse, nsteps,
                                   y_pulse_out, y_out, avg_out, stats_pulse, stats_relax, false);
}

extern "C" int stb_integrate_dynamics_dev(stb_system* S, int32_t first, int32_t n, int64_t n_tp,
                                          const double* t_pulse_dev, int64_t n_t, const double* t_sol_dev,
                                          const double* y0_planes_dev, double rtol, double atol,
                                          int64_t nsteps_pulse, int64_t nsteps, double* y_pulse_out_dev,
                                          double* y_out_dev, double* avg_out_dev, stb_stats* stats_pulse,
                                          stb_stats* stats_relax) {
    return integrate_dynamics_core(S, first, n, n_tp, t_pulse_dev, n_t, t_sol_dev, y0_planes_dev, rtol, atol,
                                   nsteps_pulse, nsteps, y_pulse_out_dev, y_out_dev, avg_out_dev, stats_pulse,
                                   stats_relax, true);
}

extern "C" int stb_set_experiment(stb_system* S, int32_t kind, int64_t n_e, const double* t_exp,
                                  const double* y_exp, double tail_mean) {
    if (!S) return fail(STB_EINVAL, "null system");
    const int kinds = S->hL.s[0] + S->hL.s[1];
    if (kind < 0 || kind >= kinds) return fail(STB_EINVAL, "state kind out of range");
    if (n_e < 0 || (n_e > 0 && (!t_exp || !y_exp))) return fail(STB_EINVAL, "bad experimental curve");
    if (S->exp_t.empty()) {
        S->exp_t.assign(kinds, {});
        S->exp_y.assign(kinds, {});
        S->exp_tail.assign(kinds, 0.0);
    }
    S->exp_t[kind].assign(t_exp, t_exp + n_e);
    S->exp_y[kind].assign(y_exp, y_exp + n_e);
    S->exp_tail[kind] = tail_mean;
    S->exp_dirty = true;
    return STB_OK;
}

extern "C" int stb_objective(stb_system* S, int32_t n, double* errors) {
    if (!S || !errors) return fail(STB_EINVAL, "null argument");
    if (S->exp_t.empty()) return fail(STB_ESTATE, "no experimental curve set (stb_set_experiment)");
    if (!S->last_avg || n != S->last_dyn_n)
        return fail(STB_ESTATE, "stb_objective evaluates the curves of the last stb_integrate_dynamics call");
    CK(cudaSetDevice(S->device));
    const int kinds = S->hL.s[0] + S->hL.s[1];
    if (S->exp_dirty) {
        std::vector<int64_t> off(kinds + 1, 0);
        for (int k = 0; k < kinds; ++k) off[k + 1] = off[k] + (int64_t)S->exp_t[k].size();
        const int64_t tot = std::max<int64_t>(off[kinds], 1);
        if (tot > S->exp_cap) {
            if (S->d_exp_t) pool_free(S->d_exp_t);
            if (S->d_exp_y) pool_free(S->d_exp_y);
            S->d_exp_t = S->d_exp_y = nullptr;
            CK(pool_malloc(&S->d_exp_t, sizeof(double) * tot));
            CK(pool_malloc(&S->d_exp_y, sizeof(double) * tot));
            S->exp_cap = tot;
        }
        if (!S->d_exp_off) CK(pool_malloc(&S->d_exp_off, sizeof(int64_t) * (kinds + 1)));
        if (!S->d_exp_tail) CK(pool_malloc(&S->d_exp_tail, sizeof(double) * kinds));
        for (int k = 0; k < kinds; ++k) {
            const size_t nb = sizeof(double) * S->exp_t[k].size();
            if (!nb) continue;
            CK(copy_on(S->stream, S->d_exp_t + off[k], S->exp_t[k].data(), nb, cudaMemcpyHostToDevice));
            CK(copy_on(S->stream, S->d_exp_y + off[k], S->exp_y[k].data(), nb, cudaMemcpyHostToDevice));
        }
        CK(copy_on(S->stream, S->d_exp_off, off.data(), sizeof(int64_t) * (kinds + 1), cudaMemcpyHostToDevice));
        CK(copy_on(S->stream, S->d_exp_tail, S->exp_tail.data(), sizeof(double) * kinds, cudaMemcpyHostToDevice));
        S->exp_dirty = false;
    }
    if ((int64_t)n * kinds > S->obj_cap) {
        if (S->d_obj) pool_free(S->d_obj);
        S->d_obj = nullptr;
        CK(pool_malloc(&S->d_obj, sizeof(double) * (size_t)n * kinds));
        S->obj_cap = (int64_t)n * kinds;
    }
    g_launches++;
    k_objective<<<dim3((unsigned)kinds, (unsigned)n), 256, 0, S->stream>>>(
        S->last_avg, S->last_ts, S->last_dyn_nt, kinds, S->d_exp_t, S->d_exp_y, S->d_exp_off, S->d_exp_tail,
        S->d_obj);
    CK(cudaGetLastError());
    CK(copy_on(S->stream, errors, S->d_obj, sizeof(double) * (size_t)n * kinds, cudaMemcpyDeviceToHost));
    return STB_OK;
}

extern "C" int stb_set_option(stb_system* S, const char* name, double value) {
    if (!S || !name) return fail(STB_EINVAL, "null argument");
    if (!strcmp(name, "sm_budget")) {
        S->sm_budget = (int)std::max(1.0, std::min((double)S->sm_count, value));
        return STB_OK;
    }
    if (!strcmp(name, "avg_cumulative")) {
        S->avg_cumulative = value != 0.0;
        return STB_OK;
    }
    return fail(STB_EINVAL, std::string("unknown option ") + name);
}

extern "C" int stb_get_option(const stb_system* S, const char* name, double* value) {
    if (!S || !name || !value) return fail(STB_EINVAL, "null argument");
    if (!strcmp(name, "sm_budget")) *value = S->sm_budget;
    else if (!strcmp(name, "quad")) *value = S->hL.quad;
    else if (!strcmp(name, "pc_group")) *value = S->hL.quad ? S->hL.pcg : 1;
    else if (!strcmp(name, "cell_lists")) *value = S->built_with_cells ? 1.0 : 0.0;
    else return fail(STB_EINVAL, std::string("unknown option ") + name);
    return STB_OK;
}

extern "C" int stb_last_timing_dynamics(const stb_system* S, float* ms_pulse, float* ms_relax) {
    if (!S) return fail(STB_EINVAL, "null system");
    if (ms_pulse) *ms_pulse = S->last_ms_pulse;
    if (ms_relax) *ms_relax = S->last_ms_relax;
    return STB_OK;
}

extern "C" void* stb_stream(const stb_system* S) { return S ? (void*)S->stream : nullptr; }


// ---- cooperative (three-ion) process --------------------------------------------------------
extern "C" int stb_set_coop(stb_system* S, const double* dist_array, const double* positions,
                            const double* cell, const int32_t* idx6, int32_t mult, double strength,
                            double d_max_coop) {
    if (!S || !idx6 || (!dist_array && !positions)) return fail(STB_EINVAL, "bad arguments");
    CK(cudaSetDevice(S->device));
    CK(cudaStreamSynchronize(S->stream));
    DevLattice& L = S->hL;
    // drop a previous cooperative process: lists, field slots and terms
    for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 4; ++k) { pool_free(S->d_coop[r][k]); S->d_coop[r][k] = nullptr; }
    S->h_coop_N.clear(); S->h_coop_rows.clear(); S->h_coop_w.clear();
    if (L.n_coop == 0) {
        for (int T = 0; T < 2; ++T) { S->n_terms_pair[T] = L.terms[T].n_terms; S->nf_pair[T] = L.nf[T]; }
    }
    for (int T = 0; T < 2; ++T) { L.terms[T].n_terms = S->n_terms_pair[T]; L.nf[T] = S->nf_pair[T]; }
    L.n_coop = 0;
    const int n = (int)S->n_ions;
    const int ii = idx6[0], ki = idx6[1], li = idx6[2], if_ = idx6[3], kf = idx6[4], lf = idx6[5];
    const bool active = stb_active_strength(strength) && S->N > 5 && S->n[0] >= 2 && S->n[1] >= 1;
    if (active) {
        if (ii >= S->s[0] || if_ >= S->s[0] || ki >= S->s[0] || kf >= S->s[0] || li >= S->s[1] || lf >= S->s[1] ||
            ii < 0 || if_ < 0 || ki < 0 || kf < 0 || li < 0 || lf < 0)
            return fail(STB_EINVAL, "The number of A or S states is lower than required by a process.");
        if (L.n_proc >= STB_MAX_PROC) return fail(STB_EINVAL, "too many processes");
        if (L.nf[0] + 2 > 2 * STB_MAXC || L.nf[1] + 1 > 2 * STB_MAXC ||
            L.terms[0].n_terms + 2 > STB_MAXT || L.terms[1].n_terms + 1 > STB_MAXT)
            return fail(STB_EINVAL, "too many distinct field slots");
        Geo G;
        memset(&G, 0, sizeof(G));
        G.n = n;
        G.dist = dist_array;
        G.pos = dist_array ? nullptr : positions;
        if (!dist_array && cell) {
            G.periodic = 1;
            memcpy(G.cell, cell, sizeof(double) * 9);
            if (!invert3(G.cell, G.inv)) return fail(STB_EINVAL, "singular cell");
        }
        // reference order (precalculate.py:471-480): acceptors in lattice order, pairs (i < k) of
        // their sensitizer partners in combination order, both closer than d_max_coop
        std::vector<int> s_ions;
        for (int q = 0; q < n; ++q) if (S->type[q] == 0) s_ions.push_back(q);
        const bool cut = d_max_coop > 0.0 && std::isfinite(d_max_coop);
        std::vector<int64_t> cnt[3];
        cnt[0].assign((size_t)S->n[0] + 1, 0); cnt[1].assign((size_t)S->n[0] + 1, 0); cnt[2].assign((size_t)S->n[1] + 1, 0);
        struct Tri { int i, k, l; double w; };
        std::vector<Tri> tri;
        std::vector<double> dl(s_ions.size());
        std::vector<int> keep;
        for (int l = 0; l < n; ++l) {
            if (S->type[l] != 1) continue;
            keep.clear();
            for (size_t a = 0; a < s_ions.size(); ++a) {
                dl[a] = geo_dist(G, l, s_ions[a]);
                if (dl[a] != 0.0) keep.push_back((int)a);       // np.nonzero(dist), lattice.py:335
            }
            for (size_t x = 0; x < keep.size(); ++x)
                for (size_t y = x + 1; y < keep.size(); ++y) {
                    const int a = keep[x], b = keep[y];
                    const double dli = dl[a], dlk = dl[b];
                    if (cut && !(std::max(dli, dlk) < d_max_coop)) continue;
                    const double dik = geo_dist(G, s_ions[a], s_ions[b]);
                    const double w = 1.0 / std::pow(dli * dlk, (double)mult) + 1.0 / std::pow(dli * dik, (double)mult) +
                                     1.0 / std::pow(dik * dlk, (double)mult);
                    tri.push_back({s_ions[a], s_ions[b], l, strength * w});
                }
        }
        const size_t Mc = tri.size();
        S->h_coop_N.resize(Mc * 3); S->h_coop_rows.resize(Mc * 6); S->h_coop_w.resize(Mc);
        for (size_t r = 0; r < Mc; ++r) {
            const Tri& t = tri[r];
            const int gi = S->gs_ref[t.i], gk = S->gs_ref[t.k], gl = S->gs_ref[t.l];
            S->h_coop_N[3 * r] = gi + ii; S->h_coop_N[3 * r + 1] = gk + ki; S->h_coop_N[3 * r + 2] = gl + li;
            int32_t* rw = &S->h_coop_rows[6 * r];
            rw[0] = gi + ii; rw[1] = gi + if_; rw[2] = gk + ki; rw[3] = gk + kf; rw[4] = gl + li; rw[5] = gl + lf;
            S->h_coop_w[r] = t.w;
            cnt[0][S->local_idx[t.i] + 1]++; cnt[1][S->local_idx[t.k] + 1]++; cnt[2][S->local_idx[t.l] + 1]++;
        }
        if (Mc > 0) {
            for (int r = 0; r < 3; ++r)
                for (size_t q = 1; q < cnt[r].size(); ++q) cnt[r][q] += cnt[r][q - 1];
            std::vector<uint32_t> p1[3], p2[3];
            std::vector<double> ww[3];
            std::vector<int64_t> fill[3];
            for (int r = 0; r < 3; ++r) { p1[r].resize(Mc); p2[r].resize(Mc); ww[r].resize(Mc); fill[r] = cnt[r]; }
            for (size_t r = 0; r < Mc; ++r) {
                const Tri& t = tri[r];
                const uint32_t li_ = S->local_idx[t.i], lk_ = S->local_idx[t.k], ll_ = S->local_idx[t.l];
                int64_t e = fill[0][li_]++; p1[0][e] = lk_; p2[0][e] = ll_; ww[0][e] = t.w;   // owner i: (k, l)
                e = fill[1][lk_]++; p1[1][e] = li_; p2[1][e] = ll_; ww[1][e] = t.w;           // owner k: (i, l)
                e = fill[2][ll_]++; p1[2][e] = li_; p2[2][e] = lk_; ww[2][e] = t.w;           // owner l: (i, k)
            }
            for (int r = 0; r < 3; ++r) {
                CK(upload(S->stream, (int64_t**)&S->d_coop[r][0], cnt[r]));
                CK(upload(S->stream, (uint32_t**)&S->d_coop[r][1], p1[r]));
                CK(upload(S->stream, (uint32_t**)&S->d_coop[r][2], p2[r]));
                CK(upload(S->stream, (double**)&S->d_coop[r][3], ww[r]));
            }
            // roles: 0 = first sensitizer (ii -> if; partners k [S, ki], l [A, li]),
            //        1 = second sensitizer (ki -> kf; partners i [S, ii], l [A, li]),
            //        2 = acceptor (li -> lf; partners i [S, ii], k [S, ki])
            const int role_T[3] = {0, 0, 1}, role_a[3] = {ii, ki, li}, role_b[3] = {if_, kf, lf};
            const int T1[3] = {0, 0, 0}, c1[3] = {ki, ii, ii}, T2[3] = {1, 1, 0}, c2[3] = {li, li, ki};
            for (int r = 0; r < 3; ++r) {
                CoopList& C = L.coop[r];
                C.row_ptr = (const int64_t*)S->d_coop[r][0];
                C.p1 = (const uint32_t*)S->d_coop[r][1];
                C.p2 = (const uint32_t*)S->d_coop[r][2];
                C.w = (const double*)S->d_coop[r][3];
                C.T = role_T[r]; C.T1 = T1[r]; C.col1 = c1[r]; C.T2 = T2[r]; C.col2 = c2[r];
                C.a = role_a[r]; C.b = role_b[r];
                const int T = role_T[r];
                C.fslot = L.nf[T]++;
                TermTable& tt = L.terms[T];
                const int k = tt.n_terms++;
                tt.fslot[k] = (short)C.fslot; tt.a[k] = (short)C.a; tt.b[k] = (short)C.b;
                tt.proc[k] = (short)L.n_proc;          // reserved strength slot, always 1 (w carries the strength)
            }
            L.n_coop = 3;
        }
    }
    L.foff[0] = 0;
    L.foff[1] = L.nf[0] * L.n[0];
    L.NF = L.nf[0] * L.n[0] + L.nf[1] * L.n[1];
    CK(copy_on(S->stream, S->dL, &L, sizeof(DevLattice), cudaMemcpyHostToDevice));
    S->scratch_n = 0;       // the field scratch depends on NF
    return STB_OK;
}

extern "C" int64_t stb_coop_count(const stb_system* S) { return S ? (int64_t)S->h_coop_w.size() : -1; }

extern "C" int stb_export_coop(const stb_system* S, uint32_t* coop_N_indices, int32_t* coop_rows, double* coop_w) {
    if (!S) return fail(STB_EINVAL, "null system");
    const size_t Mc = S->h_coop_w.size();
    if (coop_N_indices && Mc) memcpy(coop_N_indices, S->h_coop_N.data(), sizeof(uint32_t) * 3 * Mc);
    if (coop_rows && Mc) memcpy(coop_rows, S->h_coop_rows.data(), sizeof(int32_t) * 6 * Mc);
    if (coop_w && Mc) memcpy(coop_w, S->h_coop_w.data(), sizeof(double) * Mc);
    return STB_OK;
}

extern "C" int64_t stb_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int stb_last_timing(const stb_system* S, float* ms, int64_t* launches) {
    if (!S) return fail(STB_EINVAL, "null system");
    if (ms) *ms = S->last_ms;
    if (launches) *launches = S->last_launches;
    return STB_OK;
}

// Development aid: per-phase clock64 cycles of the last stb_integrate call, summed over its
// systems: [0] RHS passes, [1] Arnoldi operator passes, [2] Krylov dot products, [3] group
// barriers / reductions, [4] vector updates, [5] Givens, [7] rest; out[8] = CTAs per group,
// out[9] = groups.
extern "C" int stb_last_profile(const stb_system* S, double out[10]) {
    if (!S) return fail(STB_EINVAL, "null system");
    for (int k = 0; k < 10; ++k) out[k] = 0.0;
    for (size_t i = 0; i < S->last_prof.size(); ++i) out[i % 8] += (double)S->last_prof[i];
    out[8] = S->last_G;
    out[9] = S->last_groups;
    return STB_OK;
}
