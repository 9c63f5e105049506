// Latency micro-benchmarks behind the design of the cluster integrator (DEVELOPMENT TOOL).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu && /tmp/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define STB_MAXS 12
__device__ __forceinline__ double fast_rcp(double x) {
    double r = (double)__frcp_rn((float)x);
    r = r * fma(-x, r, 2.0);
    r = r * fma(-x, r, 2.0);
    return r;
}
template <int SM, bool FAST>
__device__ __forceinline__ void gj(double (&a)[STB_MAXS], int s, int W, int gl) {
#pragma unroll
    for (int col = 0; col < SM; ++col) {
        if (col < s) {
            const double piv = __shfl_sync(0xffffffffu, a[col], col, W);
            const double inv = FAST ? fast_rcp(piv) : 1.0 / piv;
            const bool mine = gl == col;
            const double f = a[col];
#pragma unroll
            for (int k = 0; k < SM; ++k) {
                if (k != col) {
                    const double pr = __shfl_sync(0xffffffffu, a[k], col, W) * inv;
                    a[k] = mine ? pr : fma(-f, pr, a[k]);
                }
            }
            a[col] = mine ? inv : -f * inv;
        }
    }
}

__global__ void k_lat(long long* out, double* sink, int s) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-3 * i;
    __syncthreads();
    long long t0, t1;
    double x = 1.0 + lane * 1e-9, y = 1.000001;
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
    // 2. dependent 64-bit shuffle chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0);
    // 3. dependent LDS.64 chain (address depends on the loaded value)
    int idx = lane;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) {
        const double v = sm[idx];
        idx = (idx + (v > 100.0 ? 1 : 2)) & 4095;
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0);
    x += idx;
    // 4. IEEE division chain / fast reciprocal chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = 1.0 / (x + 1.5);
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = (t1 - t0);
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = fast_rcp(x + 1.5);
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = (t1 - t0);
    // 5. sqrt / rsqrt chains
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = sqrt(x + 1.5);
    t1 = clock64();
    if (threadIdx.x == 0) out[5] = (t1 - t0);
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = rsqrt(x + 1.5);
    t1 = clock64();
    if (threadIdx.x == 0) out[6] = (t1 - t0);
    // 6. Gauss-Jordan of an s x s block, 16 lanes per block, one warp active per SMSP
    double a[STB_MAXS];
    const int gl = lane & 15;
#pragma unroll
    for (int k = 0; k < STB_MAXS; ++k) a[k] = (k == gl ? 2.0 : 0.0) + 0.01 * sm[(gl * 12 + k) & 4095];
    __syncthreads();
    for (int rep = 0; rep < 3; ++rep) {        // cold, then warm instruction cache
        t0 = clock64();
        gj<STB_MAXS, false>(a, s, 16, gl);
        t1 = clock64();
        if (threadIdx.x == 0) out[11 + rep] = (t1 - t0);
#pragma unroll
        for (int k = 0; k < STB_MAXS; ++k) a[k] = (k == gl ? 2.0 : 0.0) + 0.01 * a[k];
    }
    if (threadIdx.x == 0) out[7] = out[11];
#pragma unroll
    for (int k = 0; k < STB_MAXS; ++k) a[k] = (k == gl ? 2.0 : 0.0) + 0.01 * a[k];
    __syncthreads();
    t0 = clock64();
    gj<STB_MAXS, true>(a, s, 16, gl);
    t1 = clock64();
    if (threadIdx.x == 0) out[8] = (t1 - t0);
    // 7. __syncthreads with the full CTA
    __syncthreads();
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) out[9] = (t1 - t0);
    // 8. LDS.128 throughput: every warp streams 64 x 16 B per lane
    double2 acc = make_double2(0.0, 0.0);
    __syncthreads();
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 64; ++i) {
        const double2 v = *reinterpret_cast<const double2*>(sm + ((i * 64 + lane * 2) & 4094));
        acc.x += v.x;
        acc.y += v.y;
    }
    t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[10] = (t1 - t0);
    double r = x + acc.x + acc.y;
#pragma unroll
    for (int k = 0; k < STB_MAXS; ++k) r += a[k];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __cluster_dims__(4, 1, 1) k_cluster(long long* out) {
    long long t0 = clock64();
    for (int i = 0; i < 64; ++i)
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// L2 -> SM streaming: 8 CTAs of 512 threads read a 13 MB dense block, (a) as the cluster integrator
// does (a warp = 8 rows x 4 lanes x 16 B, rows `pitch` apart), (b) fully coalesced (512 B per warp
// request); DEPTH independent loads per lane are issued before the first use.
__device__ __forceinline__ double2 ld_nc2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
template <int DEPTH, bool ROWS>
__global__ void k_stream(const double* g, int n_rows, int pitch, double* sink, long long* cyc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows_per_cta = n_rows / gridDim.x;
    double acc = 0.0;
    long long t0 = clock64();
    if (ROWS) {
        const int rr = lane & 7, kk = lane >> 3;
        const int nst = pitch / 8;                       // steps of 8 doubles along a row
        // (row group, step) work split evenly over the warps like the integrator does
        const int n_groups = rows_per_cta / 8;
        const long long total = (long long)n_groups * nst;
        long long w0 = total * warp / 16, w1 = total * (warp + 1) / 16;
        while (w0 < w1) {
            const int grp = (int)(w0 / nst), s0 = (int)(w0 % nst);
            const int s1 = (int)min((long long)nst, s0 + (w1 - w0));
            const double* gr = g + ((size_t)(blockIdx.x * rows_per_cta + grp * 8 + rr)) * pitch + 2 * kk;
            int st = s0;
            for (; st + DEPTH <= s1; st += DEPTH) {
                double2 v[DEPTH];
#pragma unroll
                for (int u = 0; u < DEPTH; ++u) v[u] = ld_nc2(gr + 8 * (st + u));
#pragma unroll
                for (int u = 0; u < DEPTH; ++u) acc += v[u].x + v[u].y;
            }
            for (; st < s1; ++st) { double2 v = ld_nc2(gr + 8 * st); acc += v.x + v.y; }
            w0 += s1 - s0;
        }
    } else {
        const size_t per_cta = (size_t)rows_per_cta * pitch;
        const double* base = g + blockIdx.x * per_cta;
        const size_t per_warp = per_cta / 16;
        const double* p = base + warp * per_warp + 2 * lane;
        const size_t n_it = per_warp / 64;               // 64 doubles = 512 B per warp request
        size_t it = 0;
        for (; it + DEPTH <= n_it; it += DEPTH) {
            double2 v[DEPTH];
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) v[u] = ld_nc2(p + 64 * (it + u));
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) acc += v[u].x + v[u].y;
        }
    }
    long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int DEPTH, bool ROWS>
static void run_stream(const double* g, int n_rows, int pitch, double* sink, long long* d_out, const char* name) {
    for (int rep = 0; rep < 3; ++rep) k_stream<DEPTH, ROWS><<<8, 512>>>(g, n_rows, pitch, sink, d_out);
    long long c;
    cudaMemcpy(&c, d_out, sizeof(c), cudaMemcpyDeviceToHost);
    const double bytes = (double)n_rows * pitch * 8 / 8;       // per CTA
    printf("L2 stream %-34s depth %2d: %8.0f cycles per pass, %5.1f B/clk/SM\n", name, DEPTH, (double)c, bytes / c);
}

int main() {
    long long* d_out;
    double* d_sink;
    cudaMalloc(&d_out, 16 * sizeof(long long));
    cudaMalloc(&d_sink, 512 * sizeof(double));
    const char* names[] = {"DFMA dependent (per op)", "SHFL.64 dependent (per op)", "LDS.64 dependent (per op)",
                           "IEEE 1/x dependent (per op, incl. add)", "fast rcp dependent (per op, incl. add)",
                           "sqrt dependent (per op)", "rsqrt dependent (per op)", "Gauss-Jordan 9x9 IEEE (total)",
                           "Gauss-Jordan 9x9 fast rcp (total)", "__syncthreads (per barrier)",
                           "LDS.128 stream, all warps (per load)"};
    const double div[] = {256, 256, 256, 64, 64, 64, 64, 1, 1, 64, 64};
    for (int threads : {32, 128, 512}) {
        cudaMemset(d_out, 0, 16 * sizeof(long long));
        k_lat<<<1, threads, 4096 * 8>>>(d_out, d_sink, 9);
        long long h[16];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("--- %d threads per CTA\n", threads);
        for (int k = 0; k < 11; ++k) printf("%-45s %8.1f cycles\n", names[k], h[k] / div[k]);
        printf("Gauss-Jordan 9x9 IEEE, calls 1..3: %lld %lld %lld cycles\n", h[11], h[12], h[13]);
    }
    {
        const int n_rows = 1280, pitch = 1280;          // 13.1 MB of doubles: L2-resident
        double* g;
        double* sink2;
        cudaMalloc(&g, sizeof(double) * (size_t)n_rows * pitch);
        cudaMalloc(&sink2, sizeof(double) * 8 * 512);
        cudaMemset(g, 0, sizeof(double) * (size_t)n_rows * pitch);
        run_stream<1, true>(g, n_rows, pitch, sink2, d_out, "8 rows x 64 B per warp request");
        run_stream<4, true>(g, n_rows, pitch, sink2, d_out, "8 rows x 64 B per warp request");
        run_stream<16, true>(g, n_rows, pitch, sink2, d_out, "8 rows x 64 B per warp request");
        run_stream<1, false>(g, n_rows, pitch, sink2, d_out, "512 B contiguous per warp request");
        run_stream<4, false>(g, n_rows, pitch, sink2, d_out, "512 B contiguous per warp request");
        run_stream<16, false>(g, n_rows, pitch, sink2, d_out, "512 B contiguous per warp request");
    }
    k_cluster<<<4, 512>>>(d_out);
    long long h0;
    cudaMemcpy(&h0, d_out, sizeof(h0), cudaMemcpyDeviceToHost);
    printf("cluster barrier (4 CTAs x 512 threads)        %8.1f cycles\n", h0 / 64.0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
