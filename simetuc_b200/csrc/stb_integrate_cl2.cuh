// Cluster integrator, second layout of the inner loops (same algorithm, constants and data layout
// as stb_integrate_cl.cuh, which stays as the fallback for lattices whose CTAs own more ion states
// than threads).  Replaces odesolver._solve_ode (odesolver.py:85-137) for small dense lattices.
//
// What changes against k_integrate_cl_t: the per-step latency chain.
//   * ion-major thread ownership: a thread owns ONE state of ONE ion for the whole launch, the
//     states of an ion sit in adjacent lanes of one warp.  The per-ion work of an iteration (rate
//     terms, rows of M = I - cJ_qq, block-Jacobi apply) runs in registers and warp shuffles: no
//     shared-memory round trip and no CTA barrier between the pair pass and the reductions;
//   * the Krylov dot products are accumulated by the owners (8 values per butterfly of 9 shuffles),
//     summed per CTA by one warp and exchanged through distributed shared memory; orthogonalisation,
//     normalisation and the exchange of the next direction follow in registers.  A Krylov
//     iteration = pair pass (2 CTA barriers) + 1 CTA barrier + 2 cluster barriers;
//   * reductions return their totals in registers of every thread (no broadcast barrier);
//   * the scalar step prologue (step clipping, c = h / alpha, preconditioner age) is evaluated
//     redundantly by every thread instead of by one thread between two barriers.
#pragma once
#include "stb_integrate_cl.cuh"

#define CL2_CH 8           // dot products per butterfly / exchange chunk

// threads a CTA needs for the ion-major ownership (host: feasibility; kernel: mapping)
__host__ __device__ inline int cl2_threads(const DevLattice& L, int cnt0, int cnt1) {
    const int w0 = cl_pow2ceil(L.s[0] > 1 ? L.s[0] : 1), w1 = cl_pow2ceil(L.s[1] > 1 ? L.s[1] : 1);
    return ((cnt0 * w0 + 31) & ~31) + cnt1 * w1;
}

// Sum of 8 per-lane values over the warp in 9 shuffles: on return the lanes with (lane & 3) == 0
// hold the total of value ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1).
__device__ __forceinline__ double cl2_warp_reduce8(const double (&v)[8], int lane) {
    const unsigned full = 0xffffffffu;
    double u[4], t[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double send = b4 ? v[q] : v[q + 4];
        const double keep = b4 ? v[q + 4] : v[q];
        u[q] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const double send = b3 ? u[q] : u[q + 2];
        const double keep = b3 ? u[q + 2] : u[q];
        t[q] = keep + __shfl_xor_sync(full, send, 8);
    }
    const double send = b2 ? t[0] : t[1];
    double r = (b2 ? t[1] : t[0]) + __shfl_xor_sync(full, send, 4);
    r += __shfl_xor_sync(full, r, 2);
    r += __shfl_xor_sync(full, r, 1);
    return r;
}


// Pair pass of one CTA (see k_integrate_cl_t): fields of the own rows from the gathered vector at
// offset zb of the dynamic shared array; zb / lin_at as in cl_coop_fields.  Not inlined on purpose.
template <bool COOP>
__device__ __noinline__ long long cl2_fields_pass(ClShared* Sp, int zb, double mult, int lin_at, long long tick) {
    __builtin_assume(__isShared(Sp));
    STB_DYN_SMEM(dynb);
    ClShared& S = *Sp;
    const DevLattice& L = S.L;
    const ClLayout& Y = S.Y;
    const ClConst& K = S.K;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rr = lane & 7, kk = lane >> 3;
    const int my_ga = S.wb_g[warp], my_ta = S.wb_t[warp];
    const int my_gb = S.wb_g[warp + 1], my_tb = S.wb_t[warp + 1];
    const int NRG = K.nrg[0] + K.nrg[1];
    double* F = dynb + Y.F;
    double* Pp = dynb + Y.Pp;
#define CL2PROF(k)                                           \
    if (tid == 0) {                                          \
        const long long now_ = clock64();                    \
        S.prof[k] += now_ - tick;                            \
        tick = now_;                                         \
    }
    int g = my_ga, t0 = my_ta, piece = 0;
    while (g < NRG && (g < my_gb || (g == my_gb && t0 < my_tb))) {
        const int T = g >= K.nrg[0];
        const int gi = T ? g - K.nrg[0] : g;
        const int tend = g == my_gb ? my_tb : K.ST[T];
        const int i = min(gi * CL_RG + rr, K.cnt[T] - 1);
#pragma unroll 1
        for (int U = 0; U < 2; ++U) {
            const int lo = U ? K.stp[T][0] : 0;
            const int s0 = max(t0, lo) - lo, s1 = min(tend, lo + K.stp[T][U]) - lo;
            if (s0 >= s1) continue;
            const int ns = L.blk[T][U].n_slots;
            int dst, dpitch;                 // offsets into the dynamic shared array
            if (s0 == 0 && s1 == K.stp[T][U]) {
                dst = Y.F + Y.foff[T] + (U ? L.blk[T][0].n_slots : 0) * Y.cntp[T] + gi * CL_RG + rr;
                dpitch = Y.cntp[T];
            } else {
                dst = Y.Pp + ((warp * 2 + (piece ? 1 : 0)) * Y.nsm) * CL_RG + rr;
                dpitch = CL_RG;
            }
            if (K.gres_on[T][U])
                cl_piece_ns<true>(ns, K.goff[T][U], nullptr, K.gpitch[T][U], i, zb + Y.zoff[U],
                                  Y.ldz[U], L.blk[T][U].slot_col, Y.nkp[U], kk, s0, s1, mult, dst, dpitch);
            else
                cl_piece_ns<false>(ns, 0, K.gsrc[T][U], K.gpitch[T][U], i, zb + Y.zoff[U],
                                   Y.ldz[U], L.blk[T][U].slot_col, Y.nkp[U], kk, s0, s1, mult, dst, dpitch);
            ++piece;
        }
        ++g;
        t0 = 0;
    }
    CL2PROF(0)
    __syncthreads();
    CL2PROF(1)
    // segments shared by several warps: sum the pieces in warp order
    const int nc = S.n_comb;
    if (nc > 0) {
        for (int idx = tid; idx < nc; idx += STB_NT) {
            const unsigned e = S.comb[idx];
            const int w0 = (e >> 16) & 15, w1 = (e >> 20) & 15, sl0 = (e >> 24) & 1;
            const int q = (int)((e >> 25) & 7) * CL_RG + (int)(e >> 28);
            double acc = Pp[((w0 * 2 + sl0) * Y.nsm) * CL_RG + q];
            for (int w = w0 + 1; w <= w1; ++w) acc += Pp[(w * 2 * Y.nsm) * CL_RG + q];
            F[e & 0xffffu] = acc;
        }
        __syncthreads();
    }
    if (COOP) {
        cl_coop_fields(L, Y, K.r_lo, K.cnt, zb, mult, lin_at, warp, lane);
        __syncthreads();
    }
    CL2PROF(2)
#undef CL2PROF
    return tick;
}

template <bool COOP>
__global__ void __launch_bounds__(STB_NT, 1) k_integrate_cl2_t(ClArgs A) {
    STB_DYN_SMEM(dyn);
    __shared__ ClShared S;
    __shared__ double s_dred[2][STB_NW][CL2_CH];     // per-warp partials of the 8 reduced values
    DevLattice& L = S.L;
    SysParams& P = S.P;
    Ctl& C = S.C;
    ClLayout& Y = S.Y;

    const int G = A.G;
    const int rank = G > 1 ? (int)cl_rank() : 0;
    const int cid = blockIdx.x / G;
    const int n_clusters = gridDim.x / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tid = threadIdx.x;

    for (int i = tid; i < (int)(sizeof(DevLattice) / 4); i += STB_NT) ((int*)&L)[i] = ((const int*)A.L)[i];
    __syncthreads();
    if (tid == 0) cl_layout(L, G, A.resmask, A.v_smem, A.d_smem, Y, 1);
    __syncthreads();
    const int N = L.N;

    long long tick = clock64();
    if (tid == 0) for (int k = 0; k < 24; ++k) S.prof[k] = 0;
#define CLPROF(k)                                            \
    if (tid == 0) {                                          \
        const long long now_ = clock64();                    \
        S.prof[k] += now_ - tick;                            \
        tick = now_;                                         \
    }
    // ---- ownership ---------------------------------------------------------------------------
    ClConst& K = S.K;
    int (&r_lo)[2] = K.r_lo;
    int (&cnt)[2] = K.cnt;
    int (&ob)[2] = K.ob;
    int (&nrg)[2] = K.nrg;
    int (&stp)[2][2] = K.stp;
    int (&cst)[2][2] = K.cst;
    int (&ST)[2] = K.ST;
    int (&WT)[2] = K.WT;
    int (&Wt)[2] = K.Wt;
    const double* (&gsrc)[2][2] = K.gsrc;
    long long (&gpitch)[2][2] = K.gpitch;
    int (&gres_on)[2][2] = K.gres_on;
    double (&GAMMA)[6] = K.GAMMA;
    double (&ALPHA)[6] = K.ALPHA;
    double (&ERRC)[6] = K.ERRC;
    if (tid == 0) {
        for (int T = 0; T < 2; ++T) {
            stb_own_rows(L, T, rank, G, r_lo[T], cnt[T]);
            nrg[T] = (cnt[T] + CL_RG - 1) / CL_RG;
        }
        ob[0] = 0;
        ob[1] = L.s[0] * cnt[0];
        // work list of the pair pass: row groups of type 0 first; a row group of type T has
        // nst[0] (if block [T][0] is used) + nst[1] steps; a step of block [T][U] costs about
        // 2 + 3 * n_slots instruction slots
        for (int T = 0; T < 2; ++T) {
            for (int U = 0; U < 2; ++U) {
                const bool used = L.blk[T][U].n_slots > 0 && L.n[U] > 0;
                stp[T][U] = used ? Y.nst[U] : 0;
                cst[T][U] = A.cst_base + A.cst_slot * L.blk[T][U].n_slots;
                gsrc[T][U] = L.blk[T][U].g[0] ? L.blk[T][U].g[0] + (int64_t)r_lo[T] * L.ld[U] : nullptr;
                gpitch[T][U] = L.ld[U];
                gres_on[T][U] = Y.gres_off[T][U] >= 0 && used && cnt[T] > 0;
            }
            ST[T] = stp[T][0] + stp[T][1];
            // every piece also pays a fixed price (addresses, cross-lane reduction, store): CL_PIECE_FIX
            // pseudo-steps in front of each segment keep warps with three short pieces from
            // finishing long after warps with one long piece
            WT[T] = (stp[T][0] ? (stp[T][0] + A.piece_fix) * cst[T][0] : 0) +
                    (stp[T][1] ? (stp[T][1] + A.piece_fix) * cst[T][1] : 0);
            Wt[T] = cl_pow2ceil(max(L.s[T], 1));
        }
        const double KAPPA[6] = {0.0, -0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0.0};
        GAMMA[0] = 0.0;
        for (int k = 1; k <= 5; ++k) GAMMA[k] = GAMMA[k - 1] + 1.0 / k;
        for (int k = 0; k <= 5; ++k) {
            ALPHA[k] = (1.0 - KAPPA[k]) * GAMMA[k];
            ERRC[k] = KAPPA[k] * GAMMA[k] + 1.0 / (k + 1);
        }
    }
    __syncthreads();
    const int n_own = ob[1] + L.s[1] * cnt[1];
    const int NRG = nrg[0] + nrg[1];

    double* zy = dyn + Y.zy;
    double* zv = dyn + Y.zv;
    double* dd = dyn + Y.dd;
    double* psi = dyn + Y.psi;
    double* scl = dyn + Y.scl;
    double* fv = dyn + Y.fv;
    double* tw = dyn + Y.tw;
    double* kv = dyn + Y.kv;
    double* Minv = dyn + Y.minv;
    double* Mcur = dyn + Y.mcur;
    const bool lagpc = (A.opts & CL_OPT_LAGPC) != 0;
    const bool guess = (A.opts & CL_OPT_GUESS) != 0;
    double* F = dyn + Y.F;
    double* Pp = dyn + Y.Pp;
    double* part = dyn + Y.part;
    int* own_ez = (int*)(dyn + Y.info);         // index into the gathered (padded plane) vector
    int* own_inf = own_ez + Y.PO;               // T << 30 | a << 24 | own row i
    int* own_ref = own_inf + Y.PO;              // index in the reference (ion-major) layout
    double* wk = A.work + ((int64_t)cid * G + rank) * A.work_stride;
    const int PO = Y.PO;
    // Krylov basis [KRYLOV+1][PO]: the first nvs vectors in shared memory (a linear solve rarely
    // needs more), the rest spilled to the work area; difference array D [8][PO]
    const int nvs = A.v_smem;
    auto Vrow = [&](int i) -> double* {
        return i < nvs ? dyn + Y.V + i * PO : wk + (int64_t)(i - nvs) * PO;
    };
    double* D = A.d_smem ? dyn + Y.D : wk + (int64_t)(CL_KRYLOV + 1) * PO;

#ifndef STB_CL_NO_ZEROFILL
    // start from a defined state: nothing below may depend on what an earlier kernel left in the
    // SM's shared memory (padding columns are multiplied by g = 0, and 0 * NaN is not 0)
    for (int i = tid; i < Y.total; i += STB_NT) dyn[i] = 0.0;
#ifdef STB_TEST_ALIGN128           // test builds only: the 32 B gap the alignment request opens below dyn
    if (tid < 4) dyn[-1 - tid] = 0.0;
#endif
    __syncthreads();
#endif
    for (int i = tid; i < 2 * Y.NZ; i += STB_NT) zy[i] = 0.0;     // zy and zv are contiguous
    for (int i = tid; i < Y.part - Y.F; i += STB_NT) F[i] = 0.0;  // unused field slots and pieces of idle warps stay 0

    // ---- rate terms by state -----------------------------------------------------------------
    if (tid == 0) {
        for (int T = 0; T < 2; ++T) {
            const TermTable& tt = L.terms[T];
            int n = 0;
            for (int a = 0; a <= L.s[T]; ++a) {
                S.tl_ptr[T][a] = (short)n;
                if (a == L.s[T]) break;
                for (int t = 0; t < tt.n_terms; ++t) {
                    if (tt.a[t] == a) S.tl_ent[T][n++] = (short)t;
                    if (tt.b[t] == a) S.tl_ent[T][n++] = (short)(t | 0x4000);
                }
            }
        }
    }

    // ---- work list of the pair pass: (row group, partner step), equal cost per warp ----------
    if (tid <= STB_NW) {
        const int64_t Ctot = (int64_t)nrg[0] * WT[0] + (int64_t)nrg[1] * WT[1];
        int64_t x = Ctot * tid / STB_NW;
        int g = 0, t = 0;
        if (tid == STB_NW || Ctot == 0) {
            g = NRG;
        } else {
            int T = 0;
            if (WT[0] > 0 && x < (int64_t)nrg[0] * WT[0]) {
                g = (int)(x / WT[0]);
                x -= (int64_t)g * WT[0];
            } else {
                x -= (int64_t)nrg[0] * WT[0];
                T = 1;
                const int gi = WT[1] > 0 ? (int)(x / WT[1]) : 0;
                x -= (int64_t)gi * WT[1];
                g = nrg[0] + gi;
            }
            const int64_t c0 = stp[T][0] ? (int64_t)(stp[T][0] + A.piece_fix) * cst[T][0] : 0;
            if (x < c0) t = max(0, (int)(x / cst[T][0]) - A.piece_fix);
            else t = stp[T][0] + (stp[T][1] ? min(stp[T][1], max(0, (int)((x - c0) / cst[T][1]) - A.piece_fix)) : 0);
        }
        S.wb_g[tid] = g;
        S.wb_t[tid] = t;
    }
    __syncthreads();
    // linear step positions of the warp boundaries; segments (row group, block) cut by a boundary
    // get one combine entry per (field slot, row)
    if (tid == 0) {
        K.rgpos[0] = 0;
        K.rgpos[1] = nrg[0] * ST[0];
        S.n_comb = 0;
    }
    __syncthreads();
    if (tid <= STB_NW) {
        const int g = S.wb_g[tid];
        const int T = g >= nrg[0];
        S.wpos[tid] = g >= NRG ? nrg[0] * ST[0] + nrg[1] * ST[1]
                               : K.rgpos[T] + (g - (T ? nrg[0] : 0)) * ST[T] + S.wb_t[tid];
    }
    __syncthreads();
    for (int sg = tid; sg < 2 * NRG; sg += STB_NT) {
        const int g = sg >> 1, U = sg & 1;
        const int T = g >= nrg[0];
        const int gi = T ? g - nrg[0] : g;
        const int a0 = K.rgpos[T] + gi * ST[T] + (U ? stp[T][0] : 0), b0 = a0 + stp[T][U];
        if (b0 <= a0) continue;
        int w0 = -1, w1 = -1;
        for (int w = 0; w < STB_NW; ++w)
            if (max(S.wpos[w], a0) < min(S.wpos[w + 1], b0)) {
                if (w0 < 0) w0 = w;
                w1 = w;
            }
        if (w0 < 0 || w1 == w0) continue;              // one warp: written to F directly
        // slot of w0's piece: 0 if this segment holds the first step of w0's range
        const int sl0 = S.wpos[w0] >= a0 ? 0 : 1;
        const int ns = L.blk[T][U].n_slots;
        const int base = atomicAdd(&S.n_comb, ns * CL_RG);
        const int sb = U ? L.blk[T][0].n_slots : 0;     // first field slot of the block in F
        for (int q = 0; q < ns * CL_RG; ++q) {
            const int sl = q / CL_RG, r = q % CL_RG;
            const unsigned dst = (unsigned)(Y.foff[T] + (sb + sl) * Y.cntp[T] + gi * CL_RG + r);
            if (base + q < CL_MAXCOMB)
                S.comb[base + q] = dst | (w0 << 16) | (w1 << 20) | (sl0 << 24) | (sl << 25) | (r << 28);
        }
    }
    __syncthreads();
    const int my_ga = S.wb_g[warp], my_ta = S.wb_t[warp];
    const int my_gb = S.wb_g[warp + 1], my_tb = S.wb_t[warp + 1];

    // ---- pair rows of the own ions -> shared memory, [pair][own row][2] -----------------------
    {
        double* gres = dyn + Y.gres;
        for (int T = 0; T < 2; ++T)
            for (int U = 0; U < 2; ++U) {
                if (!gres_on[T][U]) continue;
                double* dst = gres + Y.gres_off[T][U];
                const int cp = Y.cntm[T], kcols = Y.nst[U] * 8;       // padded partner columns
                const double* src = gsrc[T][U];
                for (int idx = tid; idx < cp * kcols; idx += STB_NT) {
                    const int i = idx / kcols, k = idx % kcols;
                    double v = 0.0;
                    if (i < cnt[T] && k < L.n[U]) v = src[(int64_t)i * L.ld[U] + k];
                    dst[((int64_t)(k >> 1) * cp + i) * 2 + (k & 1)] = v;
                }
            }
        __syncthreads();
        if (tid == 0)
            for (int T = 0; T < 2; ++T)
                for (int U = 0; U < 2; ++U)
                    if (gres_on[T][U]) {
                        K.goff[T][U] = Y.gres + Y.gres_off[T][U];
                        gpitch[T][U] = Y.cntm[T];
                    }
    }
    __syncthreads();


    // ---- ion-major ownership: one thread = one state of one own ion ------------------------------
    const unsigned FULLM = 0xffffffffu;
    const int R0 = (cnt[0] * Wt[0] + 31) & ~31;            // type-0 region, whole warps
    const int mT = tid < R0 ? 0 : 1;
    const int midx = tid - (mT ? R0 : 0);
    const int mW = Wt[mT];
    const int mi = midx / mW, ma = midx % mW;
    const int ms = L.s[mT];                                // warp-uniform: a warp holds one ion type
    const bool valid = mi < cnt[mT] && ma < ms;
    const int lane0 = lane & ~(mW - 1);                    // first lane of the ion's lane group
    const int cmT = Y.cntm[mT];
    const int k = valid ? ob[mT] + ma * cnt[mT] + mi : 0;  // own-compact index (storage layout unchanged)
    const int pl = valid ? L.off[mT] + ma * L.n[mT] + r_lo[mT] + mi : 0;       // plane index
    const int ez = valid ? Y.zoff[mT] + ma * Y.ldz[mT] + r_lo[mT] + mi : 0;    // gathered-vector index
    const int refi = valid ? L.ref_of_plane[pl] : 0;                           // reference-layout index
    const bool lagT = lagpc && ms > 2;
    // warps without any own element skip the owner phases (they take part in the pair passes, the
    // Gauss-Jordan teams and every barrier): the owner phases are bound by issue slots
    const bool wown = (warp * 32 < R0) ? (warp * 32 < cnt[0] * Wt[0]) : (warp * 32 - R0 < cnt[1] * Wt[1]);
    const int GIV = STB_NT - 32;                           // the thread that keeps the Hessenberg / Givens state
    for (int i = tid; i < 2 * STB_NW * CL2_CH; i += STB_NT) (&s_dred[0][0][0])[i] = 0.0;
    __syncthreads();

    // every CTA of the cluster must be running before its shared memory is written remotely
    if (G > 1) cl_barrier();
    // remote bases of the exchange buffers
    uint32_t rb_zy[CL_MAXG], rb_zv[CL_MAXG], rb_part[CL_MAXG];
#pragma unroll
    for (int r = 0; r < CL_MAXG; ++r) {
        const uint32_t rr = r < G ? r : 0;
        if (G > 1) {
            rb_zy[r] = cl_mapa(cl_smem_u32(zy), rr);
            rb_zv[r] = cl_mapa(cl_smem_u32(zv), rr);
            rb_part[r] = cl_mapa(cl_smem_u32(part), rr);
        } else {
            rb_zy[r] = rb_zv[r] = rb_part[r] = 0;
        }
    }
    int pbuf = 0, qbuf = 0;

    auto sync_all = [&]() {
        if (G > 1) cl_barrier();
        else __syncthreads();
    };
    // own element -> every CTA of the cluster
    auto push = [&](const uint32_t (&rb)[CL_MAXG], double* local, int e, double v) {
        if (G > 1) {
#pragma unroll
            for (int r = 0; r < CL_MAXG; ++r)
                if (r < G) cl_st_f64(rb[r] + (uint32_t)e * 8u, v);
        } else {
            local[e] = v;
        }
    };
    // CTA partial `mine` of value idx (one thread per value) -> slot of this rank in every CTA
    auto push_part = [&](int idx, double mine) {
        const int slot = (pbuf * G + rank) * CL_KRED + idx;
        if (G > 1) {
#pragma unroll
            for (int r = 0; r < CL_MAXG; ++r)
                if (r < G) cl_st_f64(rb_part[r] + (uint32_t)slot * 8u, mine);
        } else {
            part[slot] = mine;
        }
    };
    // butterfly of the 8 per-thread values, then the CTA sum of values [0, nb) goes to slot c0 + ..
    // of the exchange buffer.  Contains one CTA barrier; the cluster barrier is the caller's.
    const int vi_lane = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    auto cta_partials = [&](const double (&v)[8], int nb, int c0) {
        if (wown) {
            const double r = cl2_warp_reduce8(v, lane);
            if ((lane & 3) == 0) s_dred[qbuf][warp][vi_lane] = r;
        }
        CLPROF(21)
        __syncthreads();
        CLPROF(22)
        if (tid < nb) {
            double mine = 0.0;
#pragma unroll
            for (int w = 0; w < STB_NW; ++w) mine += s_dred[qbuf][w][tid];
            push_part(c0 + tid, mine);
        }
        qbuf ^= 1;
    };
    // cluster total of exchanged value idx, summed in rank order (identical in every thread)
    auto total_of = [&](int idx) -> double {
        const double* pp = part + pbuf * G * CL_KRED + idx;
        double acc = 0.0;
        for (int r = 0; r < G; ++r) acc += pp[r * CL_KRED];
        return acc;
    };
    // all-reduce of K <= 8 per-thread values: totals in tot[] of every thread, no trailing barrier
    auto allreduce8 = [&](int K, const double (&v)[8], double (&tot)[8]) {
        cta_partials(v, K, 0);
        sync_all();
#pragma unroll
        for (int q = 0; q < 8; ++q) tot[q] = q < K ? total_of(q) : 0.0;
        pbuf ^= 1;
    };

    // ---- pair pass: F[slot][row] = mult * sum_k g(row, k) z[col_slot][k] for the own rows: ONE
    // out-of-line instance for the four call sites (the kernel is bound by instruction fetch:
    // its per-iteration code must stay inside the 32 KB instruction cache) ----------------------------
    auto fields_pass = [&](int zb, double mult, int lin_at) {
        tick = cl2_fields_pass<COOP>(&S, zb, mult, lin_at, tick);
    };

    // rate terms of element (T, a, i): sum_t C_t y_q[a_t] F[slot_t][i] ([b_t == a] - [a_t == a])
    auto term_sum = [&](int T, int a, int i, const double* y /*gathered*/) -> double {
        const TermTable& tt = L.terms[T];
        const double* yq = y + Y.zoff[T] + r_lo[T] + i;
        const double* Fq = F + Y.foff[T] + i;
        double acc = 0.0;
        for (int n = S.tl_ptr[T][a]; n < S.tl_ptr[T][a + 1]; ++n) {
            const int e = S.tl_ent[T][n];
            const int t = e & 0x3fff;
            const double rate = P.C[tt.proc[t]] * yq[tt.a[t] * Y.ldz[T]] * Fq[tt.fslot[t] * Y.cntp[T]];
            acc += (e & 0x4000) ? rate : -rate;
        }
        return acc;
    };
    auto lin_sum = [&](int T, int a, int i, const double* y) -> double {
        const double* yq = y + Y.zoff[T] + r_lo[T] + i;
        const double* lr = P.lin[T] + a * L.s[T];
        double acc = 0.0;
        for (int b = 0; b < L.s[T]; ++b) acc = fma(lr[b], yq[b * Y.ldz[T]], acc);
        return acc;
    };
    // x_a = sum_b M[a][b] r_b over the states of the own ion: r travels by warp shuffle
    auto ion_matvec = [&](const double* M, double r) -> double {
        const double* Mp = M + Y.moff[mT] + (ma * ms) * cmT + mi;
        double acc = 0.0;
        if (!wown) return acc;
        for (int b = 0; b < ms; ++b) {
            const double rb = __shfl_sync(FULLM, r, lane0 + b);
            if (valid) acc = fma(Mp[b * cmT], rb, acc);
        }
        return acc;
    };

    // block-Jacobi: M = I - c (Lin_T + sum_t C_t F[slot_t] (e_b - e_a) e_a^T) per own ion.  Every
    // thread builds the row of its own state (into Mcur when the inverses are lagged); 1- and
    // 2-state ions are inverted in closed form by their first thread; larger blocks by in-place
    // Gauss-Jordan when `refresh` (see k_integrate_cl_t).
    auto build_blocks = [&](double c, bool refresh) {
        if (valid) {
            const TermTable& tt = L.terms[mT];
            const double* lin = P.lin[mT];
            const double* Fq = F + Y.foff[mT] + mi;
            if (ms <= 2) {
                if (ma == 0) {
                    double m00 = 1.0 - c * lin[0], m01 = 0.0, m10 = 0.0, m11 = 1.0;
                    if (ms == 2) { m01 = -c * lin[1]; m10 = -c * lin[2]; m11 = 1.0 - c * lin[3]; }
                    for (int t = 0; t < tt.n_terms; ++t) {
                        const double kq = c * P.C[tt.proc[t]] * Fq[tt.fslot[t] * Y.cntp[mT]];
                        if (tt.a[t] == 0) { m00 += kq; m10 -= kq; }       // column a_t: row a_t +, row b_t -
                        else { m11 += kq; m01 -= kq; }
                    }
                    double* Mp = Minv + Y.moff[mT] + mi;
                    if (ms == 1) {
                        Mp[0] = m00 != 0.0 ? 1.0 / m00 : 0.0;
                    } else {
                        const double det = m00 * m11 - m01 * m10;
                        const double id = det != 0.0 ? 1.0 / det : 0.0;
                        Mp[0] = m11 * id;
                        Mp[cmT] = -m01 * id;
                        Mp[2 * cmT] = -m10 * id;
                        Mp[3 * cmT] = m00 * id;
                    }
                }
            } else {
                double* Mp = (lagpc ? Mcur : Minv) + Y.moff[mT] + (ma * ms) * cmT + mi;
                const double* lr = lin + ma * ms;
                for (int b = 0; b < ms; ++b) Mp[b * cmT] = (b == ma ? 1.0 : 0.0) - c * lr[b];
                for (int n = S.tl_ptr[mT][ma]; n < S.tl_ptr[mT][ma + 1]; ++n) {
                    const int e = S.tl_ent[mT][n];
                    const int t = e & 0x3fff;
                    const double kq = c * P.C[tt.proc[t]] * Fq[tt.fslot[t] * Y.cntp[mT]];
                    Mp[tt.a[t] * cmT] += (e & 0x4000) ? -kq : kq;     // row b_t: -, row a_t: +
                }
                if (lagpc && refresh) {
                    double* Mi = Minv + Y.moff[mT] + (ma * ms) * cmT + mi;
                    for (int b = 0; b < ms; ++b) Mi[b * cmT] = Mp[b * cmT];
                }
            }
        }
        if (!refresh) {
            __syncwarp();          // the closed-form inverse is read by the other lane of the ion
            return;
        }
        __syncthreads();
        for (int T = 0; T < 2; ++T) {
            const int s = L.s[T], cm = Y.cntm[T];
            if (s <= 2 || cnt[T] <= 0) continue;
            const int ss = s * s, tl = tid & 63;
            int ei[3], ej[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int e = tl + 64 * u;
                ei[u] = e / s;
                ej[u] = e - ei[u] * s;
            }
            for (int ion0 = 0; ion0 < cnt[T]; ion0 += STB_NT / 64) {
                const int ion = ion0 + (tid >> 6);
                const bool act = ion < cnt[T];
                double* Mq = Minv + Y.moff[T] + (act ? ion : 0);
                for (int p = 0; p < s; ++p) {
                    double nv[3];
                    if (act) {
                        const double piv = Mq[(p * s + p) * cm];
                        const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            const int e = tl + 64 * u;
                            if (e < ss) {
                                const int i = ei[u], j = ej[u];
                                const double aip = Mq[(i * s + p) * cm] * inv, apj = Mq[(p * s + j) * cm];
                                nv[u] = i == p ? (j == p ? inv : apj * inv)
                                               : (j == p ? -aip : fma(-aip, apj, Mq[e * cm]));
                            }
                        }
                    }
                    __syncthreads();
                    if (act) {
#pragma unroll
                        for (int u = 0; u < 3; ++u)
                            if (tl + 64 * u < ss) Mq[(tl + 64 * u) * cm] = nv[u];
                    }
                    __syncthreads();
                }
            }
        }
    };

    auto apply_change_D = [&](int order) {
        if (valid) {
            double dv[6], nd[6];
            for (int j = 0; j <= order; ++j) dv[j] = D[j * PO + k];
            for (int i = 0; i <= order; ++i) {
                double acc = 0.0;
                for (int j = 0; j <= order; ++j) acc = fma(C.RU[j * 6 + i], dv[j], acc);
                nd[i] = acc;
            }
            for (int i = 0; i <= order; ++i) D[i * PO + k] = nd[i];
        }
    };

    const double sqrtN = sqrt((double)N);
    const double rtol = A.rtol, atol = A.atol;
    const double EPS = 2.220446049250313e-16;

    // =================================== systems ============================================
    while (true) {
        // next system of the batch (dynamic: solves differ in their number of steps)
        __syncthreads();
        if (rank == 0 && tid == 0) {
            const int nxt = atomicAdd(A.next_sys, 1);
            if (G > 1) {
                for (int r = 0; r < G; ++r) cl_st_u32(cl_mapa(cl_smem_u32(&S.cur_sys), r), (uint32_t)nxt);
            } else {
                S.cur_sys = nxt;
            }
        }
        sync_all();
        const int sysl = S.cur_sys;
        if (sysl >= A.n_sys) break;
        const int sys = A.first + sysl;
        load_sys_params(L, A.blocks, A.strengths, sys, A.pulse, P);
        const double* y0 = A.y0 + (int64_t)sysl * N;
        double* y_out = A.y_out + (int64_t)sysl * A.n_t * N;
        if (tid == 0) {
            C.t = A.t[0];
            C.t_end = A.t[A.n_t - 1];
            C.order = 1;
            C.n_equal = 0;
            C.status = 0;
            C.k_out = 1;
            C.steps_since_out = 0;
            C.newton_tol = fmax(10.0 * EPS / rtol, fmin(0.03, sqrt(rtol)));
            C.tol_lin = A.lin_rms * sqrtN;
            C.crate = 1.0;
            C.since_verify = 0;
            C.gfac = 0.0;
            C.pc_c = 0.0;
            C.pc_age = -1;
            C.pc_refresh = 1;
            C.n_iter = 1;
            memset(&C.st, 0, sizeof(stb_stats));
        }
        int64_t k_out_l = 1;                                   // output cursor, uniform over the cluster
        double tq_l = A.n_t > 1 ? A.t[1] : INFINITY;
        double tq_n = A.n_t > 2 ? A.t[2] : INFINITY;           // one output time ahead: its load latency stays off the step
        // own element in registers: scale, its reciprocal, psi, the Newton correction
        double my_scl = 1.0, my_iscl = 1.0, my_psi = 0.0, my_dd = 0.0;
        // D0 = y0, row 0 of the output = y0 verbatim (odesolver.py:107)
        double yq0 = 0.0;
        if (valid) {
            yq0 = y0[pl];
            D[k] = yq0;
            for (int j = 1; j < 8; ++j) D[j * PO + k] = 0.0;
            y_out[refi] = yq0;
            push(rb_zy, zy, ez, yq0);
        }
        sync_all();
        fields_pass(Y.zy, 1.0, -1);
        // ---- initial step (Hairer-Wanner, as scipy select_initial_step) -----------------------
        {
            const double t_beg = A.t[0], t_fin = A.t[A.n_t - 1];
            double f0 = 0.0;
            double v8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, tot[8];
            if (valid) {
                f0 = lin_sum(mT, ma, mi, zy) + term_sum(mT, ma, mi, zy);
                my_scl = atol + rtol * fabs(yq0);
                const double a = yq0 / my_scl, b = f0 / my_scl;
                v8[0] = a * a;
                v8[1] = b * b;
                if (!isfinite(f0)) v8[2] = 1.0;
            }
            allreduce8(3, v8, tot);
            const double d0 = sqrt(tot[0] / N), d1 = sqrt(tot[1] / N);
            const double bad = tot[2];
            const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            if (valid) push(rb_zv, zv, ez, yq0 + h0 * f0);
            sync_all();
            fields_pass(Y.zv, 1.0, -1);
            v8[0] = v8[1] = v8[2] = 0.0;
            if (valid) {
                const double f1 = lin_sum(mT, ma, mi, zv) + term_sum(mT, ma, mi, zv);
                const double a = (f1 - f0) / my_scl;
                v8[0] = a * a;
            }
            allreduce8(1, v8, tot);
            const double d2 = sqrt(tot[0] / N) / h0;
            const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3)
                                                           : sqrt(0.01 / fmax(d1, d2));
            const double hini = fmin(fmin(100.0 * h0, h1), t_fin - t_beg);
            if (tid == 0) {
                C.st.n_rhs += 2;
                C.h = hini;
                if (bad != 0.0 || !isfinite(hini) || !(hini > 0.0)) C.status = 3;
            }
            if (valid) D[PO + k] = f0 * hini;
        }

        // =============================== time stepping =====================================
        while (true) {
            __syncthreads();
            CLPROF(14)
            if (C.status != 0 || k_out_l >= A.n_t) break;
            // scalar prologue, evaluated by every thread from the same shared values
            double h = C.h;
            const double t = C.t, t_end = C.t_end;
            const int order = C.order;
            {
                const double t_try = t + h;
                if (t_try > t_end && (t_try - t_end) > 1e-14 * fabs(t_end)) {
                    // the last step is shortened to land on t_end (rare)
                    __syncthreads();
                    if (tid == 0) {
                        const double factor = (t_end - t) / h;
                        cl_compute_RU(order, factor, C.RU, K.Rs, K.Us);
                        C.h = t_end - t;
                        C.n_equal = 0;
                        C.gfac *= cl_ipow(factor, order + 1);
                    }
                    __syncthreads();
                    apply_change_D(order);
                    h = t_end - t;
                } else if (t_try > t_end) {
                    h = t_end - t;      // round-off sized overshoot: land exactly on t_end
                }
            }
            {
                int st_now = 0;
                if (h < 10.0 * EPS * fabs(t)) st_now = 2;
                if (C.steps_since_out >= A.nsteps) st_now = 1;
                if (st_now) {
                    __syncthreads();
                    if (tid == 0) C.status = st_now;
                    break;
                }
            }
            const double c = h / ALPHA[order];
            const bool pc_ref = !lagpc || C.pc_age < 0 || C.pc_age >= A.pc_maxage ||
                                fabs(c / C.pc_c - 1.0) > A.pc_drift;
            const double gf = guess ? C.gfac : 0.0;
            bool first_newton = true;
            // predictor, scale, psi; the Newton start (predictor + the previous correction rescaled
            // to this step) goes to every CTA of the cluster
            if (valid) {
                double yp = 0.0, ps = 0.0;
                for (int j = 0; j <= order; ++j) {
                    const double dj = D[j * PO + k];
                    yp += dj;
                    if (j >= 1) ps = fma(GAMMA[j], dj, ps);
                }
                my_scl = atol + rtol * fabs(yp);
                my_iscl = 1.0 / my_scl;
                my_psi = ps / ALPHA[order];
                my_dd = gf != 0.0 ? gf * D[(order + 1) * PO + k] : 0.0;
                push(rb_zy, zy, ez, yp + my_dd);
            }
            if (tid == 0) { C.newton_it = 0; C.flag = 0; C.dy_norm_old = -1.0; }
            // ---- Newton-Krylov ------------------------------------------------------------------
            while (true) {       // flag: 0 iterate, 1 converged, 2 failed
                CLPROF(15)
                sync_all();      // y_k complete in every CTA
                CLPROF(3)
                fields_pass(Y.zy, 1.0, -1);
                double fq = 0.0;
                if (valid) fq = lin_sum(mT, ma, mi, zy) + term_sum(mT, ma, mi, zy);
                CLPROF(4)
                const bool pc_now = !lagpc || (pc_ref && first_newton);
                build_blocks(c, pc_now);
                CLPROF(5)
                // x = Minv (c f - psi - d) / scale; the unscaled vector goes to zv
                double xq;
                {
                    const double xy = ion_matvec(Minv, valid ? c * fq - my_psi - my_dd : 0.0);
                    xq = valid ? xy * my_iscl : 0.0;
                    if (valid) push(rb_zv, zv, ez, xy);
                    double v8[8] = {xq * xq, (valid && !isfinite(fq)) ? 1.0 : 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    double tot[8];
                    CLPROF(6)
                    allreduce8(2, v8, tot);
                    CLPROF(7)
                    if (tid == 0) {
                        if (lagpc && pc_now) { C.pc_c = c; C.pc_age = 0; }
                        C.st.n_rhs += 1;
                        C.st.n_lin += 1;
                        C.beta = sqrt(tot[0]);
                        C.kk = 0;
                        if (tot[1] != 0.0 || !isfinite(C.beta)) C.flag = 2;
                        S.gg[0] = C.beta;
                        C.inrm = C.beta > 0.0 ? 1.0 / C.beta : 0.0;
                        if (C.flag != 2 && !(C.beta > C.tol_lin)) {
                            // residual already below the linear tolerance: dy = 0
                            int fl = newton_decide(C, 0.0);
                            if (fl == 0 && C.newton_it >= STB_NEWTON_MAXIT) fl = 2;
                            C.flag = fl;
                        }
                    }
                    __syncthreads();
                }
                // ---- GMRES: V_0 = x / beta; zv holds the direction, scaled by `zmul` ----------
                bool lin_done = !(C.beta > C.tol_lin) || C.flag == 2;
                int j = 0;
                double zmul = C.inrm;
                if (!lin_done && valid) Vrow(0)[k] = xq * zmul;
                while (!lin_done) {
                    // w = Minv (I - cJ) v / scale = V_j - c Minv k(v) / scale, k = inter-ion coupling
                    CLPROF(15)
                    fields_pass(Y.zv, zmul, Y.zy);
                    double w = 0.0;
                    {
                        double kq = valid ? term_sum(mT, ma, mi, zy) : 0.0;
                        if (lagT && valid) {
                            // lagged inverse: Minv_old (M v - c k(v)), rows of the current M in Mcur
                            const double* Mp = Mcur + Y.moff[mT] + (ma * ms) * cmT + mi;
                            const double* vq = zv + Y.zoff[mT] + r_lo[mT] + mi;
                            double mv = 0.0;
                            for (int b = 0; b < ms; ++b) mv = fma(Mp[b * cmT], vq[b * Y.ldz[mT]], mv);
                            kq = mv * zmul - c * kq;
                        }
                        const double mk = ion_matvec(Minv, kq);
                        if (valid) w = lagT ? mk * my_iscl : Vrow(j)[k] - c * mk * my_iscl;
                    }
                    CLPROF(8)
                    // h_i = <w, V_i> (i <= j), ww = <w, w>: partial products of the own element, 8
                    // values per butterfly; then orthogonalise in registers.  Returns ||w_orth||^2 by
                    // Pythagoras and leaves h in S.rv (Givens thread).
                    const int Kd = j + 2;
                    double ww = 0.0;
                    auto dots_orth = [&]() -> double {
                        for (int c0 = 0; c0 < Kd; c0 += CL2_CH) {
                            double p8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                            if (wown) {
#pragma unroll
                                for (int q = 0; q < CL2_CH; ++q) {
                                    const int idx = c0 + q;
                                    double vi = 0.0;
                                    if (valid && idx < Kd) vi = idx <= j ? Vrow(idx)[k] : w;
                                    p8[q] = w * vi;
                                }
                            }
                            CLPROF(20)
                            cta_partials(p8, min(CL2_CH, Kd - c0), c0);
                        }
                        CLPROF(9)
                        sync_all();
                        CLPROF(10)
                        ww = total_of(j + 1);
                        double hn2 = ww;
                        for (int i = 0; i <= j; ++i) {
                            const double hi = total_of(i);
                            hn2 = fma(-hi, hi, hn2);
                            if (valid) w = fma(-hi, Vrow(i)[k], w);
                            if (tid == GIV) S.rv[i] = hi;
                        }
                        pbuf ^= 1;
                        CLPROF(23)
                        return hn2;
                    };
                    double hn2 = dots_orth();
                    // Pythagorean norm of the orthogonalised vector: with ||w_orth|| > 1e-4 ||w|| the
                    // cancellation costs at most 8 of 16 digits of hn2 and the basis stays orthogonal
                    // to ~1e-12, far inside the linear tolerance; below that (a preconditioned operator
                    // this close to the identity ends the solve anyway) re-orthogonalise explicitly
                    const bool reorth = !(hn2 > 1e-8 * ww);
                    if (reorth) {
                        // cancellation: explicit norm and one re-orthogonalisation pass (DGKS)
                        if (tid == GIV)
                            for (int i = 0; i <= j; ++i) S.H[i * CL_KRYLOV + j] = S.rv[i];
                        hn2 = dots_orth();
                        if (tid == GIV)
                            for (int i = 0; i <= j; ++i) S.H[i * CL_KRYLOV + j] += S.rv[i];
                        if (!(hn2 > 1e-8 * ww)) {
                            double v8[8] = {w * w, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, tot[8];
                            allreduce8(1, v8, tot);
                            hn2 = tot[0];
                        }
                    }
                    const double ihn = hn2 > 0.0 ? rsqrt(hn2) : 0.0;
                    const double hn = hn2 * ihn;
                    // Givens update of column j, residual estimate and (when the solve closes) the
                    // triangular solve + Newton decision: one thread of the last warp, off the path of
                    // the threads that normalise and exchange the next direction
                    if (tid == GIV) {
                        C.st.n_jv += 1;
                        double* Hc = S.H + j;   // column j, stride CL_KRYLOV
                        double prev = reorth ? Hc[0] : S.rv[0];
                        for (int i = 0; i < j; ++i) {
                            const double b = reorth ? Hc[(i + 1) * CL_KRYLOV] : S.rv[i + 1];
                            Hc[i * CL_KRYLOV] = S.cs[i] * prev + S.sn[i] * b;
                            prev = -S.sn[i] * prev + S.cs[i] * b;
                        }
                        const double a = prev;
                        const double dnm = sqrt(fma(a, a, hn * hn));
                        const double idn = dnm > 0.0 ? 1.0 / dnm : 0.0;
                        const double csj = dnm > 0.0 ? a * idn : 1.0, snj = hn * idn;
                        S.cs[j] = csj;
                        S.sn[j] = snj;
                        Hc[j * CL_KRYLOV] = dnm;
                        S.rd[j] = idn;
                        S.gg[j + 1] = -snj * S.gg[j];
                        S.gg[j] = csj * S.gg[j];
                        C.kk = j + 1;
                        double beta = fabs(S.gg[j + 1]);
                        if (!(hn > 0.0) || !isfinite(hn)) beta = 0.0;
                        C.beta = beta;
                        if (!(beta > C.tol_lin) || j + 1 >= CL_KRYLOV) {
                            // x = V y with R y = g; ||x|| = ||y|| (orthonormal basis)
                            const int nk = j + 1;
                            double y2 = 0.0;
                            for (int i = nk - 1; i >= 0; --i) {
                                double acc = S.gg[i];
                                for (int kk2 = i + 1; kk2 < nk; ++kk2) acc -= S.H[i * CL_KRYLOV + kk2] * S.yk_s[kk2];
                                const double yi = acc * S.rd[i];
                                S.yk_s[i] = yi;
                                y2 = fma(yi, yi, y2);
                            }
                            int fl = newton_decide(C, sqrt(y2 / N));
                            if (fl == 0 && C.newton_it >= STB_NEWTON_MAXIT) fl = 2;
                            C.flag = fl;     // 2: do not apply the update
                        }
                    }
                    // V_{j+1} = w / ||w||; its unscaled form is the next direction
                    if (valid) {
                        const double v = w * ihn;
                        Vrow(j + 1)[k] = v;
                        push(rb_zv, zv, ez, v * my_scl);
                    }
                    zmul = 1.0;
                    CLPROF(11)
                    sync_all();
                    CLPROF(12)
                    ++j;
                    lin_done = !(C.beta > C.tol_lin) || j >= CL_KRYLOV;
                }
                const int nk = C.kk;
                const int flag = C.flag;
                // a failed iteration (diverging, or out of iterations) is discarded with the step
                if (flag != 2 && valid) {
                    double x = 0.0;
                    for (int i = 0; i < nk; ++i) x = fma(S.yk_s[i], Vrow(i)[k], x);
                    const double dy = x * my_scl;
                    const double yn = zy[ez] + dy;
                    my_dd += dy;
                    if (flag == 0) push(rb_zy, zy, ez, yn);   // next iteration needs it
                    else zy[ez] = yn;
                }
                if (flag != 0) break;
                first_newton = false;
            }
            __syncthreads();
            if (C.flag == 2) {
                // Newton / Krylov failure: halve the step (scipy: factor 0.5 with current Jacobian)
                if (tid == 0) {
                    C.st.n_newton_fail += 1;
                    cl_compute_RU(order, 0.5, C.RU, K.Rs, K.Us);
                    C.h = h * 0.5;
                    C.n_equal = 0;
                    C.steps_since_out += 1;
                    C.gfac = 0.0;
                    C.pc_age = -1;       // factor the block inverses afresh
                }
                __syncthreads();
                apply_change_D(order);
                continue;
            }
            CLPROF(13)
            const double t_new = t + h;
            // ---- error test and order selection share one reduction -------------------------------
            double err_norm, err_m2, err_p2, safety;
            {
                double v8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, tot[8];
                if (valid) {
                    my_scl = atol + rtol * fabs(zy[ez]);
                    my_iscl = 1.0 / my_scl;
                    const double er = ERRC[order] * my_dd * my_iscl;
                    v8[0] = er * er;
                    if (order > 1) {
                        const double em = ERRC[order - 1] * (D[order * PO + k] + my_dd) * my_iscl;
                        v8[1] = em * em;
                    }
                    if (order < STB_MAXORD) {
                        const double ep = ERRC[order + 1] * (my_dd - D[(order + 1) * PO + k]) * my_iscl;
                        v8[2] = ep * ep;
                    }
                }
                const int n_iter = C.n_iter;
                allreduce8(3, v8, tot);
                err_norm = sqrt(tot[0] / N);
                err_m2 = tot[1];
                err_p2 = tot[2];
                safety = 0.9 * (2 * STB_NEWTON_MAXIT + 1) / (2.0 * STB_NEWTON_MAXIT + n_iter);
                if (err_norm > 1.0 || !isfinite(err_norm)) {
                    if (tid == 0) {
                        const double factor = isfinite(err_norm)
                            ? fmax(STB_MIN_FACTOR, safety * pow(err_norm, -1.0 / (order + 1)))
                            : STB_MIN_FACTOR;
                        cl_compute_RU(order, factor, C.RU, K.Rs, K.Us);
                        C.h = h * factor;
                        C.n_equal = 0;
                        C.st.n_rejected += 1;
                        C.steps_since_out += 1;
                        C.gfac *= cl_ipow(factor, order + 1);
                    }
                    __syncthreads();
                    apply_change_D(order);
                    continue;
                }
            }
            // ---- accepted: update differences, dense output, order / step selection -------------
            // Every thread updates and interpolates its own element with no barrier in between
            // (the output cursor k_out_l / tq_l is kept identically by all threads); the scalar
            // control work runs meanwhile on the last warp.
            if (valid) {
                const double dp1 = D[(order + 1) * PO + k];
                D[(order + 2) * PO + k] = my_dd - dp1;
                D[(order + 1) * PO + k] = my_dd;
                double run = my_dd;
                for (int i = order; i >= 0; --i) {
                    run += D[i * PO + k];
                    D[i * PO + k] = run;
                }
            }
            bool wrote = false;
            if (k_out_l < A.n_t && (tq_l <= t_new * (1.0 + 1e-15) || tq_l <= t_new)) {
                const double ih = 1.0 / h;
                const double RCPJ[5] = {1.0, 0.5, 1.0 / 3.0, 0.25, 0.2};
                do {     // dense output for every requested time in (t, t_new]
                    const bool exact = tq_l >= t_new;
                    if (valid) {
                        double acc = D[k];
                        if (!exact) {
                            double pr = 1.0;
#pragma unroll
                            for (int jj = 0; jj < STB_MAXORD; ++jj) {
                                pr *= (tq_l - (t_new - h * jj)) * (ih * RCPJ[jj]);
                                if (jj < order) acc = fma(pr, D[(jj + 1) * PO + k], acc);
                            }
                        }
                        y_out[k_out_l * (int64_t)N + refi] = acc;
                    }
                    wrote = true;
                    ++k_out_l;
                    tq_l = tq_n;
                    tq_n = k_out_l + 1 < A.n_t ? A.t[k_out_l + 1] : INFINITY;
                } while (k_out_l < A.n_t && (tq_l <= t_new * (1.0 + 1e-15) || tq_l <= t_new));
            }
            if (warp == STB_NW - 1) {
                const bool need_order = (C.n_equal + 1) >= order + 1;
                int new_order = order;
                double factor = 1.0;
                if (need_order) {
                    double f = 0.0;
                    if (lane < 3) {
                        // candidate orders order-1, order, order+1 on lanes 0, 1, 2
                        const bool avail = lane == 1 || (lane == 0 ? order > 1 : order < STB_MAXORD);
                        const double en = lane == 1 ? err_norm : sqrt((lane == 0 ? err_m2 : err_p2) / N);
                        f = !avail ? 0.0 : (en > 0.0 ? pow(en, -1.0 / (order + lane)) : INFINITY);
                    }
                    const double fm = __shfl_sync(0xffffffffu, f, 0);
                    const double f0 = __shfl_sync(0xffffffffu, f, 1);
                    const double fp = __shfl_sync(0xffffffffu, f, 2);
                    int delta;
                    double best;
                    // numpy argmax picks the first maximum: order-1 wins ties, then same order
                    if (fm >= f0 && fm >= fp) { best = fm; delta = -1; }
                    else if (f0 >= fp) { best = f0; delta = 0; }
                    else { best = fp; delta = 1; }
                    new_order = order + delta;
                    factor = fmin(STB_MAX_FACTOR, safety * best);
                    cl_compute_RU_warp(new_order, factor, C.RU, K.Rs, K.Us, lane);
                }
                if (lane == 0) {
                    C.st.n_steps += 1;
                    C.steps_since_out = wrote ? 1 : C.steps_since_out + 1;
                    C.k_out = k_out_l;
                    C.st.n_out = k_out_l;
                    C.t = t_new;
                    C.flag2 = 0;
                    if (C.pc_age >= 0) C.pc_age += 1;
                    if (need_order) {
                        C.order = new_order;
                        C.h = h * factor;
                        C.n_equal = 0;
                        C.flag2 = 1;
                        C.gfac = new_order == order ? cl_ipow(factor, order + 1) : 0.0;
                    } else {
                        C.h = h;
                        C.n_equal += 1;
                        C.gfac = 1.0;
                    }
                    C.st.last_order = C.order;
                    C.st.last_h = C.h;
                }
            }
            __syncthreads();
            if (C.flag2 == 1) apply_change_D(C.order);
        }
        __syncthreads();
        CLPROF(14)
        if (tid == 0 && rank == 0 && cid == 0 && A.prof) {
            S.prof[16] += 1;                 // systems this cluster solved; counters of all of them
            S.prof[17] += C.st.n_steps;
            S.prof[18] += C.st.n_rhs;
            S.prof[19] += C.st.n_jv;
            for (int kk2 = 0; kk2 < 24; ++kk2) A.prof[kk2] = S.prof[kk2];
        }
        if (tid == 0 && rank == 0) {
            C.st.status = C.status;
            C.st.t_reached = C.t;
            C.st.n_out = C.k_out;
            A.stats[sysl] = C.st;
        }
    }
}
