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
    k_cluster<<<4, 512>>>(d_out);
    long long h0;
    cudaMemcpy(&h0, d_out, sizeof(h0), cudaMemcpyDeviceToHost);
    printf("cluster barrier (4 CTAs x 512 threads)        %8.1f cycles\n", h0 / 64.0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
