// Standalone RHS / Jv kernels (C ABI: stb_rhs, stb_jv) and layout conversion.
// One warp per slice (R rows), grid-stride over slices, blockIdx.y = system of the batch.
#pragma once
#include "stb_coop.cuh"

struct OpArgs {
    const DevLattice* L;
    const double* blocks;
    const double* strengths;
    int first;       // first system
    int pulse;
    const double* y; // [n][N] planes
    const double* v; // [n][N] planes (Jv)
    const double* Fy;// [n][NF]  (Jv: fields at y, from a previous RHS pass)
    double* out;     // [n][N]
    double* F;       // [n][NF] (RHS: fields at y, may be null)
    int tma_kc, tma_stages, tma_diag;   // stb_tma.cuh only
};

// dy = (abs +) decay y + ET (y_i * y_j): odesolver._rate_eq / _rate_eq_pulse (odesolver.py:33-71)
#ifndef STB_RHS_MINB
#define STB_RHS_MINB 1
#endif
__global__ void __launch_bounds__(STB_NT, STB_RHS_MINB) k_rhs(OpArgs A) {
    __shared__ DevLattice L;
    __shared__ SysParams P;
    for (int i = threadIdx.x; i < (int)(sizeof(DevLattice) / 4); i += blockDim.x)
        ((int*)&L)[i] = ((const int*)A.L)[i];
    __syncthreads();
    const int b = blockIdx.y;
    load_sys_params(L, A.blocks, A.strengths, A.first + b, A.pulse, P);
    __syncthreads();
    const double* y = A.y + (int64_t)b * L.N;
    double* out = A.out + (int64_t)b * L.N;
    double* F = A.F ? A.F + (int64_t)b * L.NF : nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = L.n_slices[0] + L.n_slices[1];
    for (int task = blockIdx.x * STB_NW + warp; task < total; task += gridDim.x * STB_NW) {
        int T = task >= L.n_slices[0];
        int slice = T ? task - L.n_slices[0] : task;
        if (L.lpr >= 16) {      // warp-cooperative epilogue: lane j <-> state j
            const int W = L.lpr, gl = lane & (W - 1), s = L.s[T], nT = L.n[T];
            const double fl = slice_fields_impl<STB_UN_KERNEL, true>(L, T, slice, y, lane, 1.0, nullptr);
            const int q = slice * L.R + lane / W;
            const bool valid = q < nT;
            const int e = L.off[T] + (gl < s ? gl : 0) * nT + (valid ? q : nT - 1);
            const double yj = gl < s ? y[e] : 0.0;
            const double fj = coop_rhs(L, P, T, W, gl, yj, fl);
            if (valid && gl < s) out[e] = fj;
            if (F && valid && gl < L.nf[T]) F[L.foff[T] + gl * nT + q] = fl;
            continue;
        }
        double Fq[2 * STB_MAXC];
        slice_fields(L, T, slice, y, lane, 1.0, Fq);
        int q = slice * L.R + lane / L.lpr;
        if ((lane % L.lpr) == 0 && q < L.n[T]) {
            double yq[STB_MAXS], oq[STB_MAXS];
            load_row(L, T, q, y, yq);
            row_rhs(L, P, T, yq, Fq, oq);
            store_row(L, T, q, out, oq);
            if (F) {
                double* Fp = F + L.foff[T] + q;
                for (int k = 0; k < L.nf[T]; ++k) Fp[(int64_t)k * L.n[T]] = Fq[k];
            }
        }
    }
}

// jv = J(y) v without the dense Jacobian (odesolver.py:46-62,74-80); the fields at y come from
// the RHS pass at the same y, so one Jv streams the pair data exactly once.
__global__ void __launch_bounds__(STB_NT, STB_RHS_MINB) k_jv(OpArgs A) {
    __shared__ DevLattice L;
    __shared__ SysParams P;
    for (int i = threadIdx.x; i < (int)(sizeof(DevLattice) / 4); i += blockDim.x)
        ((int*)&L)[i] = ((const int*)A.L)[i];
    __syncthreads();
    const int b = blockIdx.y;
    load_sys_params(L, A.blocks, A.strengths, A.first + b, A.pulse, P);
    __syncthreads();
    const double* y = A.y + (int64_t)b * L.N;
    const double* v = A.v + (int64_t)b * L.N;
    const double* Fy = A.Fy + (int64_t)b * L.NF;
    double* out = A.out + (int64_t)b * L.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = L.n_slices[0] + L.n_slices[1];
    for (int task = blockIdx.x * STB_NW + warp; task < total; task += gridDim.x * STB_NW) {
        int T = task >= L.n_slices[0];
        int slice = T ? task - L.n_slices[0] : task;
        if (L.lpr >= 16) {
            const int W = L.lpr, gl = lane & (W - 1), s = L.s[T], nT = L.n[T];
            const double fvl = slice_fields_impl<STB_UN_KERNEL, true>(L, T, slice, v, lane, 1.0, nullptr);
            const int q = slice * L.R + lane / W;
            const bool valid = q < nT;
            const int qc = valid ? q : nT - 1;
            const int e = L.off[T] + (gl < s ? gl : 0) * nT + qc;
            const double yj = gl < s ? y[e] : 0.0, vj = gl < s ? v[e] : 0.0;
            const double fyl = gl < L.nf[T] ? Fy[L.foff[T] + gl * nT + qc] : 0.0;
            const double oj = coop_jv(L, P, T, W, gl, yj, vj, fyl, fvl);
            if (valid && gl < s) out[e] = oj;
            continue;
        }
        double Fv[2 * STB_MAXC];
        slice_fields(L, T, slice, v, lane, 1.0, Fv);
        int q = slice * L.R + lane / L.lpr;
        if ((lane % L.lpr) == 0 && q < L.n[T]) {
            double Fq[2 * STB_MAXC], yq[STB_MAXS], vq[STB_MAXS], oq[STB_MAXS];
            const double* Fp = Fy + L.foff[T] + q;
            for (int k = 0; k < L.nf[T]; ++k) Fq[k] = Fp[(int64_t)k * L.n[T]];
            load_row(L, T, q, y, yq);
            load_row(L, T, q, v, vq);
            row_jv(L, P, T, yq, vq, Fq, Fv, oq);
            store_row(L, T, q, out, oq);
        }
    }
}

// reference (ion-major) layout <-> plane layout, n vectors of length N
__global__ void k_ref_to_planes(const int* plane_of_ref, int N, int n, const double* ref,
                                double* planes) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < (int64_t)N * n) {
        int b = i / N, e = i % N;
        planes[(int64_t)b * N + plane_of_ref[e]] = ref[i];
    }
}

// row `y_row` (+ b * sys_stride for system b) of an output array in reference layout -> planes [n][N]:
// hands the state at the end of the pulse to the relaxation without leaving the device
__global__ void k_row_to_planes(const int* plane_of_ref, int N, int n, const double* y_row,
                                int64_t sys_stride, double* planes) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < (int64_t)N * n) {
        int b = i / N, e = i % N;
        planes[(int64_t)b * N + plane_of_ref[e]] = y_row[(int64_t)b * sys_stride + e];
    }
}

__global__ void k_planes_to_ref(const int* plane_of_ref, int N, int n, const double* planes,
                                double* ref) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < (int64_t)N * n) {
        int b = i / N, e = i % N;
        ref[i] = planes[(int64_t)b * N + plane_of_ref[e]];
    }
}

// per-state ion averages clipped at zero: Solution._calculate_avg_populations
// (simulations.py:100-128).  y_out is [n_t][N] in reference layout for one system; one block per
// (time row, system); out [n_t][sS+sA].
// clip = 0 (avg_cumulative): the raw means, for k_cumtrapz to integrate and clip afterwards.
__global__ void k_avg_populations(const DevLattice* Lg, const double* y_out, int64_t n_t,
                                  double* avg, int clip) {
    const DevLattice& L = *Lg;
    int64_t row = blockIdx.x;
    int b = blockIdx.y;
    const double* y = y_out + ((int64_t)b * n_t + row) * L.N;
    double* o = avg + ((int64_t)b * n_t + row) * (L.s[0] + L.s[1]);
    __shared__ double red[32];
    int kinds = L.s[0] + L.s[1];
    for (int k = 0; k < kinds; ++k) {
        int T = k >= L.s[0];
        int st = T ? k - L.s[0] : k;
        double acc = 0.0;
        for (int i = threadIdx.x; i < L.n[T]; i += blockDim.x)
            acc += y[L.ref_of_plane[L.off[T] + st * L.n[T] + i]];
        for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
            t = L.n[T] ? t / L.n[T] : 0.0;
            o[k] = (t > 0.0 || !clip) ? t : 0.0;
        }
        __syncthreads();
    }
}


// ---- cooperative (three-ion) energy transfer: odesolver.py:40-41 (RHS), :56-60 (Jacobian) --------
// Field of ion q in role r: Fc = sum over its partner pairs of w * z1[p1] * z2[p2]; `lin`: the
// derivative of the field in the direction v at y, sum of w * (v1 y2 + y1 v2).  One warp per ion,
// lanes stride over the ion's list (16 B per entry, coalesced), shuffle reduction: deterministic.
// out += cooperative contribution (RHS when v == null, J(y) v otherwise); planes layout,
// blockIdx.y = system of the batch.  The fused pair kernels see zero in the cooperative field
// slots, so nothing is counted twice.
__global__ void __launch_bounds__(256) k_coop_add(const DevLattice* Lp, const double* y_all, const double* v_all,
                                                  double* out_all) {
    __shared__ DevLattice L;
    for (int i = threadIdx.x; i < (int)(sizeof(DevLattice) / 4); i += blockDim.x)
        ((int*)&L)[i] = ((const int*)Lp)[i];
    __syncthreads();
    const int b = blockIdx.y;
    const double* y = y_all + (int64_t)b * L.N;
    const double* v = v_all ? v_all + (int64_t)b * L.N : nullptr;
    double* out = out_all + (int64_t)b * L.N;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n_all = L.n[0] + L.n[1];
    for (int idx = warp; idx < n_all; idx += nwarps) {
        const int T = idx >= L.n[0];
        const int q = T ? idx - L.n[0] : idx;
        for (int r = 0; r < L.n_coop; ++r) {       // the roles of one ion in turn: no two warps
            const CoopList& C = L.coop[r];          // touch the same state
            if (C.T != T) continue;
            const double* y1 = y + L.off[C.T1] + C.col1 * L.n[C.T1];
            const double* y2 = y + L.off[C.T2] + C.col2 * L.n[C.T2];
            const double Fy = coop_list_field(C, q, y1, y2, nullptr, nullptr, lane);
            double rate;
            const int ea = L.off[T] + C.a * L.n[T] + q, eb = L.off[T] + C.b * L.n[T] + q;
            if (v) {
                const double* v1 = v + L.off[C.T1] + C.col1 * L.n[C.T1];
                const double* v2 = v + L.off[C.T2] + C.col2 * L.n[C.T2];
                const double Fv = coop_list_field(C, q, y1, y2, v1, v2, lane);
                rate = fma(v[ea], Fy, y[ea] * Fv);
            } else {
                rate = y[ea] * Fy;
            }
            if (lane == 0) {
                out[ea] -= rate;
                out[eb] += rate;
            }
            __syncwarp();
        }
    }
}


// Pulsed steady state of a batch: running trapezoid of every emission curve over the solution times,
// in place (DynamicsSolution.calculate_steady_state applied to the averages, simulations.py:579-591;
// the ion average commutes with the time integral; the clip at zero does not - states whose
// population sits below atol have means of either sign - so the means come in unclipped and the
// integral is clipped, as the host does).  One thread per (system, state kind).
__global__ void k_cumtrapz(double* avg /*[n][n_t][kinds]*/, const double* t_sol, int64_t n_t, int kinds, int n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)n * kinds) return;
    const int sys = (int)(i / kinds), kind = (int)(i % kinds);
    double* a = avg + (int64_t)sys * n_t * kinds + kind;
    double prev = a[0], run = 0.0;
    a[0] = 0.0;
    for (int64_t k = 1; k < n_t; ++k) {
        const double cur = a[k * kinds];
        // same operation order as the host mirror: cumsum(0.5 * (y[k] + y[k-1]) * (t[k] - t[k-1]))
        run += 0.5 * (cur + prev) * ((t_sol[k] - t_sol[0]) - (t_sol[k - 1] - t_sol[0]));
        a[k * kinds] = run > 0.0 ? run : 0.0;
        prev = cur;
    }
}

// Optimisation objective on the device: deviation of every emission curve from its experimental
// decay curve, exactly as DynamicsSolution._get_background / _interpolate_sim_data / _calc_errors
// compute it on the host (simulations.py:425-527):
//   offset = tail_mean * max(sim)                (experimental background scaled to the simulation)
//   interp = sim interpolated linearly onto the experimental times, `offset` outside [t[0], t[-1]]
//   err    = 1 / mean(interp) * sqrt(mean((interp - exp * max(interp))^2))
// avg [n][n_t][kinds] are the curves (k_avg_populations), t_sol [n_t] the solution times (the time
// base of a solution is t_sol - t_sol[0], simulations.py:1156), exp_t / exp_y the normalised
// experimental curves of all state kinds back to back (exp_off [kinds + 1]; an empty range = no
// data = error 0).  One block per (state kind, system); fixed-order block reductions.
__device__ __forceinline__ double obj_interp(const double* avg, int64_t stride, const double* t_sol, int64_t n_t,
                                             double te, double offset) {
    const double t0 = t_sol[0];
    if (te < 0.0 || te > t_sol[n_t - 1] - t0) return offset;
    int64_t lo = 0, hi = n_t - 1;                 // t[lo] <= te <= t[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (t_sol[mid] - t0 <= te) lo = mid; else hi = mid;
    }
    const double x0 = t_sol[lo] - t0, x1 = t_sol[hi] - t0;
    const double f0 = avg[lo * stride], f1 = avg[hi * stride];
    if (te == x1) return f1;
    const double slope = (f1 - f0) / (x1 - x0);
    return slope * (te - x0) + f0;
}

template <int OP>   // 0: sum, 1: max
__device__ __forceinline__ double obj_block_reduce(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) {
        const double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = OP ? fmax(v, u) : v + u;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double acc = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) acc = OP ? fmax(acc, red[w]) : acc + red[w];
    return acc;
}

__global__ void k_objective(const double* avg, const double* t_sol, int64_t n_t, int kinds,
                            const double* exp_t, const double* exp_y, const int64_t* exp_off,
                            const double* tail_mean, double* err) {
    __shared__ double red[32];
    const int kind = blockIdx.x, sys = blockIdx.y;
    const int64_t e0 = exp_off[kind], ne = exp_off[kind + 1] - e0;
    if (ne <= 0) {
        if (threadIdx.x == 0) err[(int64_t)sys * kinds + kind] = 0.0;
        return;
    }
    const double* a = avg + (int64_t)sys * n_t * kinds + kind;
    double m = -INFINITY;
    bool any = false;
    for (int64_t k = threadIdx.x; k < n_t; k += blockDim.x) {
        const double v = a[k * kinds];
        m = fmax(m, v);
        any = any || v != 0.0;
    }
    const double sim_max = obj_block_reduce<1>(m, red);
    const double nz = obj_block_reduce<1>(any ? 1.0 : 0.0, red);
    double offset = nz != 0.0 ? tail_mean[kind] * sim_max : 0.0;
    if (!(offset > 0.0) || !isfinite(offset)) offset = 0.0;
    double s = 0.0, mx = -INFINITY;
    for (int64_t e = threadIdx.x; e < ne; e += blockDim.x) {
        const double v = obj_interp(a, kinds, t_sol, n_t, exp_t[e0 + e], offset);
        s += v;
        mx = fmax(mx, v);
    }
    const double sum_i = obj_block_reduce<0>(s, red);
    const double max_i = obj_block_reduce<1>(mx, red);
    double d = 0.0;
    for (int64_t e = threadIdx.x; e < ne; e += blockDim.x) {
        const double v = obj_interp(a, kinds, t_sol, n_t, exp_t[e0 + e], offset);
        const double r = v - exp_y[e0 + e] * max_i;
        d = fma(r, r, d);
    }
    const double dev = obj_block_reduce<0>(d, red);
    if (threadIdx.x == 0)
        err[(int64_t)sys * kinds + kind] = 1.0 / (sum_i / (double)ne) * sqrt(dev / (double)ne);
}
