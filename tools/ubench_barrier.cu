// Latency of the grid integrator's group barrier (148 CTAs x 512 threads, cooperative launch): the
// fence + atomic + acquire-spin flavour of stb_integrate.cuh, against variants (DEVELOPMENT TOOL).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_barrier tools/ubench_barrier.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// mode 0: __threadfence + atomicAdd + acquire spin + __threadfence (shipped)
// mode 1: red.release + relaxed spin (the staged flavour)
// mode 2: flags: every CTA writes its own flag line, CTA 0 gathers and releases a second flag
__global__ void k_bar(unsigned long long* counter, unsigned long long* flags, int mode, int iters, double* buf,
                      int work, long long* out) {
    const int G = gridDim.x, cta = blockIdx.x;
    unsigned long long epoch = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // some global writes before the barrier (what the fence has to publish)
        for (int i = threadIdx.x; i < work; i += blockDim.x) buf[(size_t)cta * work + i] = it + i;
        __syncthreads();
        epoch += 1;
        if (threadIdx.x == 0) {
            if (mode == 0) {
                __threadfence();
                atomicAdd(counter, 1ull);
                while (ld_acquire_u64(counter) < epoch * G) { }
                __threadfence();
            } else if (mode == 1) {
                asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(counter), "l"(1ull) : "memory");
                while (ld_relaxed_u64(counter) < epoch * G) { }
            } else {
                __threadfence();
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + 16 * cta), "l"(epoch) : "memory");
                if (cta == 0) {
                    for (int c = 0; c < G; ++c) while (ld_acquire_u64(flags + 16 * c) < epoch) { }
                    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(counter), "l"(epoch) : "memory");
                } else {
                    while (ld_acquire_u64(counter) < epoch) { }
                }
                __threadfence();
            }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[cta] = t1 - t0;
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned long long *counter, *flags;
    long long* out;
    double* buf;
    const int work_max = 4096;
    cudaMalloc(&counter, 8);
    cudaMalloc(&flags, 8 * 16 * sms);
    cudaMalloc(&out, 8 * sms);
    cudaMalloc(&buf, sizeof(double) * (size_t)sms * work_max);
    const int iters = 2000;
    for (int work : {0, 512, 4096})
        for (int mode = 0; mode < 3; ++mode) {
            cudaMemset(counter, 0, 8);
            cudaMemset(flags, 0, 8 * 16 * sms);
            int m = mode, it = iters, w = work;
            void* args[] = {&counter, &flags, &m, &it, &buf, &w, &out};
            cudaError_t e = cudaLaunchCooperativeKernel((void*)k_bar, dim3(sms), dim3(512), args, 0, 0);
            if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); return 1; }
            long long h[256];
            cudaMemcpy(h, out, 8 * sms, cudaMemcpyDeviceToHost);
            printf("work %4d doubles/CTA, mode %d: %.0f cycles per iteration (CTA 0), %.0f (last CTA)\n", work, mode,
                   (double)h[0] / iters, (double)h[sms - 1] / iters);
        }
    return 0;
}
