// Cluster integrator: the stiff integrator of stb_integrate.cuh (same BDF / Newton-Krylov /
// block-Jacobi algorithm, same constants) re-laid for SMALL lattices, where a solve is bound by
// the latency of one Krylov iteration and not by memory bandwidth.  Replaces
// odesolver._solve_ode (odesolver.py:85-137) for every system whose working set fits on a few
// SMs - the power / concentration / optimiser sweeps of BASELINE.json.
//
//   * one ODE system = one THREAD-BLOCK CLUSTER of G CTAs (G = 1, 2, 4, 8).  CTAs of a cluster
//     exchange the gathered state vector and the partial dot products through distributed
//     shared memory (st.shared::cluster) and synchronise with the hardware cluster barrier:
//     no global-memory atomics, no cooperative launch, no L2 round trip per exchange.
//   * a CTA owns a fixed set of ion rows.  Their dense pair rows are copied ONCE per launch into
//     shared memory (k-pair interleaved: one LDS.128 = two partners of one row); every hot
//     vector of the Newton-Krylov iteration lives in shared memory too.
//   * the pair pass maps a warp to 8 rows x 4 partner-pair lanes; the 16 warps take equal
//     shares of the CTA's (row group, partner step) work list, split at arbitrary points, and
//     combine partial fields deterministically (fixed warp order).
//   * per-ion work (block matvec, ET rate terms, block-Jacobi apply) is thread-per-state.
//   * one cluster barrier per reduction, two per Krylov iteration: dot products and ||w||^2
//     travel together (the new basis vector's norm follows from Pythagoras, re-orthogonalised
//     explicitly only on cancellation), ||dy|| of a Newton update is the norm of the GMRES
//     coefficients, the error test and the order selection share one reduction.
#pragma once
#include "stb_integrate.cuh"

#define CL_RG 8            // rows per row group
// GMRES basis size of the cluster integrators.  Smaller than the grid integrator's STB_KRYLOV: the
// Hessenberg / rotation arrays live in static shared memory, which the difference array needs more
// (a linear solve that does not converge in CL_KRYLOV iterations fails the Newton iteration, the
// step is halved: with the quad preconditioner solves take 1-6 iterations).
#define CL_KRYLOV 20
#define CL_KRED (CL_KRYLOV + 4)
#define CL_MAXRG 256       // row groups per CTA
#define CL_MAXG 8
#ifndef CL_STREAM_UNROLL
#define CL_STREAM_UNROLL 4     // loads in flight per lane when the pair rows are streamed from L2
#endif
constexpr int kClStreamUnroll = CL_STREAM_UNROLL;
#ifndef CL_PIECE_FIX
#define CL_PIECE_FIX 6     // fixed cost of a pair-pass piece, in steps
#endif

struct ClLayout {          // offsets in doubles into dynamic shared memory
    int NZ;                // gathered vector, padded planes
    int ldz[2], zoff[2];
    int cntm[2], cntp[2];  // max own rows per type (over the ranks), rounded up to CL_RG
    int PO;                // pitch of own-compact vectors
    int nkp[2], nst[2];    // partner pairs / warp steps (4 pairs) per partner type
    int zy, zv, dd, psi, scl, fv, tw, kv, minv, mcur, qinv, F, Pp, part, info, V, D, gres;
    int sq, nqm;               // quad lattices: states of a sensitizer quad, quads per CTA (max over ranks)
    int gres_off[2][2];    // resident pair rows of block [T][U], relative to gres (-1: streamed)
    int nfm;               // field slots per row, max over the types
    int nsm;               // field slots of one block, max over the blocks
    int foff[2], moff[2];
    int64_t total;         // doubles
};

__host__ __device__ inline int cl_pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// doubles of the shared-memory copy of a CTA's rows of block [T][U] (0: block unused)
__host__ __device__ inline int cl_block_doubles(const DevLattice& L, const ClLayout& Y, int T, int U) {
    if (L.blk[T][U].n_slots <= 0 || L.n[T] <= 0 || L.n[U] <= 0) return 0;
    return Y.nst[U] * 4 * Y.cntm[T] * 2;        // rows of the fullest rank, no row-group padding
}

// Layout shared by the host (sizing, feasibility) and the kernel.  v_smem (number of leading
// Krylov basis vectors) / d_smem / resident say
// whether the Krylov basis, the difference array and the pair rows live in shared memory.
// compact (k_integrate_cl2_t): the per-element work vectors live in registers of the owner threads
__host__ __device__ inline void cl_layout(const DevLattice& L, int G, int resmask, int v_smem,
                                          int d_smem, ClLayout& Y, int compact = 0) {
    for (int U = 0; U < 2; ++U) {
        Y.ldz[U] = (L.n[U] + 1) & ~1;
        Y.nkp[U] = Y.ldz[U] / 2;
        Y.nst[U] = (Y.nkp[U] + 3) / 4;
        Y.cntm[U] = stb_own_rows_max(L, U, G);
        Y.cntp[U] = (Y.cntm[U] + CL_RG - 1) / CL_RG * CL_RG;
    }
    Y.zoff[0] = 0;
    Y.zoff[1] = L.s[0] * Y.ldz[0];
    Y.NZ = ((Y.zoff[1] + L.s[1] * Y.ldz[1] + 8) + 1) & ~1;     // + slack for padded partner steps
    const int n_own = L.s[0] * Y.cntm[0] + L.s[1] * Y.cntm[1];
    Y.PO = (n_own + 31) & ~31;
#ifdef STB_TEST_SHIFT4              // test builds only: every array 32 B higher, base unchanged (known defect)
    int o = 4;
#else
    int o = 0;
#endif
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
    Y.zy = take(Y.NZ);
    Y.zv = take(Y.NZ);
    const int wv = compact ? 0 : Y.PO;
    Y.dd = take(wv);
    Y.psi = take(wv);
    Y.scl = take(wv);
    Y.fv = take(wv);
    Y.tw = take(wv);
    Y.kv = take(wv);
    Y.moff[0] = 0;
    Y.moff[1] = L.s[0] * L.s[0] * Y.cntm[0];
    Y.minv = take(Y.moff[1] + L.s[1] * L.s[1] * Y.cntm[1]);
    Y.mcur = take(Y.moff[1] + L.s[1] * L.s[1] * Y.cntm[1]);     // rows of I - c J_qq of the current iterate
    // quad lattices: inverse of I - cJ restricted to every sensitizer quad, [entry][quad]
    Y.sq = L.quad ? L.pcg * L.s[0] : 0;
    Y.nqm = L.quad ? Y.cntm[0] / L.pcg : 0;
    Y.qinv = take(Y.sq * Y.sq * Y.nqm);
    Y.nfm = L.nf[0] > L.nf[1] ? L.nf[0] : L.nf[1];
    if (Y.nfm < 1) Y.nfm = 1;
    Y.foff[0] = 0;
    Y.foff[1] = L.nf[0] * Y.cntp[0];
    Y.F = take(Y.foff[1] + L.nf[1] * Y.cntp[1]);
    Y.nsm = 1;
    for (int T = 0; T < 2; ++T)
        for (int U = 0; U < 2; ++U)
            if (L.blk[T][U].n_slots > Y.nsm) Y.nsm = L.blk[T][U].n_slots;
    Y.Pp = take(STB_NW * 2 * Y.nsm * CL_RG);
    Y.part = take(2 * G * CL_KRED);
    Y.info = take(compact ? 0 : Y.PO + Y.PO / 2);            // three int tables of PO entries
    Y.V = v_smem > 0 ? take(v_smem * Y.PO) : -1;       // the first v_smem basis vectors
    Y.D = d_smem ? take(8 * Y.PO) : -1;
    Y.gres = o;
    int r = 0;
    for (int T = 0; T < 2; ++T)
        for (int U = 0; U < 2; ++U) {
            Y.gres_off[T][U] = -1;
            if (!((resmask >> (2 * T + U)) & 1)) continue;
            Y.gres_off[T][U] = r;
            r += cl_block_doubles(L, Y, T, U);
        }
    o += r;
    Y.total = o;
}

// Newton starts from the previous step's correction, rescaled for a changed step size, instead of
// from the predictor alone: consecutive corrections differ by one more backward difference.
#define CL_OPT_GUESS 1
// The inverses of the larger per-ion blocks are kept across steps (refreshed when c drifts by more
// than 30 %, after CL_PC_MAXAGE steps or after a failed Newton iteration, CVODE's rule); the operator
// is then applied as Minv_old (M v - c k(v)) with the rows of the current M = I - c J_qq.
#define CL_OPT_LAGPC 2
// Quad lattices (DevLattice.quad): block preconditioner over quads of strongly coupled sensitizers
// instead of single ions - energy migration between close Yb ions is what the single-ion blocks
// miss (Krylov iterations per linear solve fall by half on the Yb-Tm sweeps).
#define CL_OPT_QUAD 4
#define CL_PC_MAXAGE 20
#define CL_PC_DRIFT 0.3

struct ClArgs {
    const DevLattice* L;
    const double* blocks;
    const double* strengths;
    int first, n_sys, pulse;
    int G;
    int resmask, v_smem, d_smem;   // bit 2T+U: rows of block [T][U] resident in shared memory
    int opts;                      // CL_OPT_* bits
    const stb_process* procs;      // [L.n_proc] the process table (quad lattices: couplings inside a quad)
    int cst_base, cst_slot;        // cost model of a pair-pass step: cst_base + cst_slot * field slots
    int piece_fix;                 // ... and the fixed cost of a piece, in steps (CL_PIECE_FIX)
    double lin_rms;                // GMRES tolerance in units of the error test (CL_LIN_RMS) ...
    int lin_tight_steps;           // ... tightened to STB_LIN_RMS for this many steps after a rejection
    // quads that do not hold the Krylov iterations down: a system that has needed more than restart_its
    // iterations per linear solve over a window of restart_steps steps within the first CL_RESTART_LAST steps
    // (what a new start may throw away is bounded) stops with status 4 and raises *restart; the
    // host repeats the launch with the 16-ion groups (null: never)
    int* restart;
    int restart_steps;
    double restart_its;
    int pc_maxage;                 // lagged block inverses: refresh after this many steps ...
    double pc_drift;               // ... or when |c / c_factored - 1| exceeds this (CL_PC_MAXAGE, CL_PC_DRIFT)
    int compact;                   // layout without the per-element work vectors (k_integrate_cl2_t)
    int64_t n_t;
    const double* t;
    const double* y0;          // [n_sys][N] planes
    double rtol, atol;
    int64_t nsteps;
    double* y_out;             // [n_sys][n_t][N] reference layout
    stb_stats* stats;
    double* work;              // [n_clusters][G][work_stride]: spilled V / D
    int64_t work_stride;
    int* next_sys;             // work counter (dynamic assignment of systems to clusters)
    const int* sys_list;       // null, or the n_list systems (local indices) this launch solves
    int n_list;
    long long* prof;           // [64] clock64 cycles per phase (cluster 0, rank 0, thread 0) + counters, per-warp pair-pass cycles and work-list positions; may be null
};

__device__ __forceinline__ double cl_ipow(double x, int n) {
    double r = 1.0;
    for (int i = 0; i < n; ++i) r *= x;
    return r;
}
__device__ __forceinline__ uint32_t cl_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t cl_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_st_f64(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void cl_st_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cl_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                     : "memory");
}
__device__ __forceinline__ uint32_t cl_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ double2 cl_ld_nc2(const double* p) {
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// One piece of the pair pass: fields of 8 rows x (partner pairs kk, kk+4, ...) over the warp
// steps [s0, s1) of one block, reduced over the 4 partner lanes of a row and stored (scaled by
// mult) at dst[slot * dpitch] by the lane kk == 0.  NS (field slots of the block) is a
// compile-time constant: no predicated slots anywhere in the piece.
//   RES: pair rows resident in shared memory, layout [pair][own row][2], addressed as offsets
//        into the dynamic shared array (the compiler then emits LDS.128, not generic loads);
//   else: global dense block, row-major, pitch ld (even), read through the non-allocating path.
// Shared memory is addressed as offsets into the dynamic shared array so that LDS / STS are
// emitted.  (Measured: a non-inlined version - one copy per (NS, RES) instead of one per call
// site - is slower: the call ABI costs more than the smaller instruction footprint saves.)
template <int NS, bool RES>
__device__ __forceinline__ void cl_piece(int g_off, const double* gb, int64_t pitch, int row, int z_off,
                                      int ldz, const int* slot_col, int nkp, int kk, int s0, int s1,
                                      double mult, int dst_off, int dpitch) {
    STB_DYN_SMEM(dynb);
    double* dst = dynb + dst_off;
    double acc[NS], acc2[NS];       // two chains per slot: the FP64 FMA latency, not its rate, binds
    int zo[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        acc[s] = 0.0;
        acc2[s] = 0.0;
        zo[s] = z_off + slot_col[s] * ldz;
    }
    if (RES) {
        const int cp = (int)pitch;
        int go = g_off + (kk * cp + row) * 2 + s0 * 8 * cp;      // [pair = st * 4 + kk][row][2]
        const int gstep = 8 * cp;
        int zk = 8 * s0 + 2 * kk;
#pragma unroll 4
        for (int st = s0; st < s1; ++st) {
            const double2 g = *reinterpret_cast<const double2*>(dynb + go);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const double2 z = *reinterpret_cast<const double2*>(dynb + zo[s] + zk);
                acc[s] = fma(g.x, z.x, acc[s]);
                acc2[s] = fma(g.y, z.y, acc2[s]);
            }
            go += gstep;
            zk += 8;
        }
    } else {
        const int kmax = nkp - 1;
        const double* gr = gb + (int64_t)row * pitch;
#pragma unroll (kClStreamUnroll)
        for (int st = s0; st < s1; ++st) {
            const int kp = st * 4 + kk;
            const double2 g = cl_ld_nc2(gr + 2 * kp);
            const int zk = 2 * min(kp, kmax);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const double2 z = *reinterpret_cast<const double2*>(dynb + zo[s] + zk);
                acc[s] = fma(g.x, z.x, acc[s]);
                acc2[s] = fma(g.y, z.y, acc2[s]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        acc[s] += acc2[s];
        acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], 8);
        acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], 16);
    }
    if (kk == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) dst[s * dpitch] = acc[s] * mult;
    }
}

template <bool RES>
__device__ __forceinline__ void cl_piece_ns(int ns, int g_off, const double* gb, int64_t pitch, int row,
                                            int z_off, int ldz, const int* slot_col, int nkp, int kk,
                                            int s0, int s1, double mult, int dst_off, int dpitch) {
#define CL_PIECE(NS_) cl_piece<NS_, RES>(g_off, gb, pitch, row, z_off, ldz, slot_col, nkp, kk, s0, s1, mult, dst_off, dpitch)
    switch (ns) {
        case 1: CL_PIECE(1); break;
        case 2: CL_PIECE(2); break;
        case 3: CL_PIECE(3); break;
        case 4: CL_PIECE(4); break;
        case 5: CL_PIECE(5); break;
        default: CL_PIECE(6); break;
    }
#undef CL_PIECE
}

#define CL_MAXCOMB ((STB_NW - 1) * STB_MAXC * CL_RG)   // a warp boundary splits at most one segment

// Cooperative (three-ion) fields of the own ions, one warp per ion and role: F[coop slot][row] =
// mult * sum w z1[p1] z2[p2], or its derivative in the direction z at the vector at lin_at
// (Krylov operator).  Not inlined: the integrator kernel is at its register limit and systems
// without a cooperative process (every shipped configuration) must not pay for this code.
__device__ __noinline__ void cl_coop_fields(const DevLattice& L, const ClLayout& Y, const int* r_lo,
                                            const int* cnt, int zb, double mult, int lin_at, int warp,
                                            int lane) {
    STB_DYN_SMEM(dynb);
    double* F = dynb + Y.F;
    for (int r = 0; r < L.n_coop; ++r) {
        const CoopList& Cl = L.coop[r];
        const int T = Cl.T;
        const int o1 = Y.zoff[Cl.T1] + Cl.col1 * Y.ldz[Cl.T1], o2 = Y.zoff[Cl.T2] + Cl.col2 * Y.ldz[Cl.T2];
        for (int i = warp; i < cnt[T]; i += STB_NW) {
            double f;
            if (lin_at >= 0)
                f = mult * coop_list_field(Cl, r_lo[T] + i, dynb + lin_at + o1, dynb + lin_at + o2,
                                           dynb + zb + o1, dynb + zb + o2, lane);
            else
                f = mult * coop_list_field(Cl, r_lo[T] + i, dynb + zb + o1, dynb + zb + o2, nullptr, nullptr, lane);
            if (lane == 0) F[Y.foff[T] + Cl.fslot * Y.cntp[T] + i] = f;
        }
    }
}

struct ClConst {           // per-CTA constants that are indexed dynamically: shared, not local memory
    int r_lo[2], cnt[2], ob[2], nrg[2];
    int stp[2][2], cst[2][2], ST[2], WT[2], Wt[2];
    const double* gsrc[2][2];  // streamed blocks: global rows of the CTA
    long long gpitch[2][2];
    int gres_on[2][2];
    int goff[2][2];            // resident blocks: offset into the dynamic shared array
    int rgpos[2];              // work-list position (steps) of the first row group of a type
    double GAMMA[6], ALPHA[6], ERRC[6], RALPHA[6];
    double Rs[36], Us[36];
};

// RU = R(order, factor) . R(order, 1) with caller-provided scratch (thread 0)
__device__ __forceinline__ void cl_compute_RU(int order, double factor, double* RU, double* R, double* U) {
    compute_R(order, factor, R);
    compute_R(order, 1.0, U);
    const int n = order + 1;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc += R[i * 6 + k] * U[k * 6 + j];
            RU[i * 6 + j] = acc;
        }
}

// The same product computed by one warp: lane j builds column j of R(order, factor), lane 8 + j
// column j of R(order, 1) (no divisions: 1/i from a table), then 36 lanes-worth of 6-term dots.
__device__ __forceinline__ void cl_compute_RU_warp(int order, double factor, double* RU, double* R,
                                                   double* U, int lane) {
    const double RCP[6] = {0.0, 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2};
    const int n = order + 1;
    if ((lane & 7) < n && lane < 16) {
        const int j = lane & 7;
        const double f = lane < 8 ? factor : 1.0;
        double* M = lane < 8 ? R : U;
        double v = 1.0;
        M[j] = 1.0;
#pragma unroll
        for (int i = 1; i < 6; ++i) {
            if (i < n) {
                v = j == 0 ? 0.0 : v * ((i - 1 - f * j) * RCP[i]);
                M[i * 6 + j] = v;
            }
        }
    }
    __syncwarp();
    for (int e = lane; e < 36; e += 32) {
        const int i = e / 6, j = e % 6;
        if (i < n && j < n) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc += R[i * 6 + k] * U[k * 6 + j];
            RU[i * 6 + j] = acc;
        }
    }
    __syncwarp();
}

struct ClShared {
    DevLattice L;
    ClConst K;
    SysParams P;
    Ctl C;
    ClLayout Y;
    double H[(CL_KRYLOV + 1) * CL_KRYLOV];
    double cs[CL_KRYLOV], sn[CL_KRYLOV], gg[CL_KRYLOV + 1], yk_s[CL_KRYLOV], rd[CL_KRYLOV];
    double red[STB_NW * 4];
    double rv[CL_KRED];
    int wb_g[STB_NW + 1], wb_t[STB_NW + 1];      // work-list boundaries of the warps
    int wpos[STB_NW + 1];                        // the same as linear step positions
    // pieces of segments shared by several warps are summed in warp order:
    // entry = dst (16) | w0 (4) | w1 (4) | slot of w0 (1) | field slot (3) | row (3)
    unsigned comb[CL_MAXCOMB];
    int n_comb;
    short tl_ptr[2][STB_MAXS + 1];               // rate terms touching state a of type T
    short tl_ent[2][2 * STB_MAXT];               // term index | sign bit (0x4000: +, i.e. b == a)
    int cur_sys;
    long long prof[24];
    long long wprof[STB_NW];                     // development aid: cycles of every warp's share of the pair passes
};

// COOP: the instance for systems with a cooperative (three-ion) process; the plain instance
// carries none of that code (it costs registers: +5 % on the C2 sweep when compiled in).
// In-place inversion of 32 x 32 blocks stored [entry][block] (block stride 1, entry stride cm): one warp per
// block, lane = one column of the stored matrix (consecutive lanes read consecutive words), eight entries at
// a time so that the loads of a chunk do not wait for its stores; no CTA barrier inside (a CTA-wide
// Gauss-Jordan pays 64 of them per block).  Out of line: the cluster kernel is instruction-cache bound and
// this runs once per preconditioner refresh.  View T[r][c] = stored[c][r]: row operations on T invert the
// stored matrix as well.  All threads of the CTA call it; ends with a CTA barrier.
__device__ __noinline__ void cl_gauss_jordan_warp32(double* base, int cm, int count, int warp, int lane) {
    for (int blk = warp; blk < count; blk += STB_NW) {
        double* Mq = base + blk;
        for (int p = 0; p < 32; ++p) {
            const double piv = Mq[(p * 32 + p) * cm];
            const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
            const double fr = Mq[(p * 32 + lane) * cm];            // T[lane][p]
            __syncwarp();
            {
                double* tp = Mq + (lane * 32 + p) * cm;             // T[p][lane]
                *tp = lane == p ? inv : *tp * inv;
            }
            __syncwarp();
            if (lane != p) {
                for (int j0 = 0; j0 < 32; j0 += 8) {
                    double pj[8], ow[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        pj[u] = Mq[((j0 + u) * 32 + p) * cm];       // T[p][j]
                        ow[u] = Mq[((j0 + u) * 32 + lane) * cm];    // T[lane][j]
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        Mq[((j0 + u) * 32 + lane) * cm] = j0 + u == p ? -fr * pj[u] : fma(-fr, pj[u], ow[u]);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
}

template <bool COOP>
__global__ void __launch_bounds__(STB_NT, 1) k_integrate_cl_t(ClArgs A) {
    STB_DYN_SMEM(dyn);
    __shared__ ClShared S;
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
    if (tid == 0) cl_layout(L, G, A.resmask, A.v_smem, A.d_smem, Y);
    __syncthreads();
    const int N = L.N;

    long long tick = clock64();
    if (tid == 0) for (int k = 0; k < 24; ++k) S.prof[k] = 0;
    if (tid < STB_NW) S.wprof[tid] = 0;
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
            K.RALPHA[k] = k ? 1.0 / ALPHA[k] : 0.0;
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
    // quad lattices: the sensitizers are preconditioned in blocks of four strongly coupled ions
    // (needs the operator in the form Minv (M v - c k(v)), i.e. the lagged mode)
    const bool quadpc = L.quad && lagpc && (A.opts & CL_OPT_QUAD) != 0 && A.procs != nullptr;
    double* Qinv = dyn + Y.qinv;
    const int pg_sh = L.pcg == 16 ? 4 : 2;      // log2 of the ions per preconditioner group
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
    for (int k = tid; k < n_own; k += STB_NT) {
        const int T = k >= ob[1];
        const int kk_ = k - ob[T];
        const int a = kk_ / cnt[T], i = kk_ % cnt[T];
        own_ez[k] = Y.zoff[T] + a * Y.ldz[T] + r_lo[T] + i;
        own_inf[k] = (T << 30) | (a << 24) | i;
        own_ref[k] = L.ref_of_plane[L.off[T] + a * L.n[T] + r_lo[T] + i];
    }

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
    int pbuf = 0;

    auto sync_all = [&]() {
        if (G > 1) cl_barrier();
        else __syncthreads();
    };
    // own element -> every CTA of the cluster
    auto push = [&](const uint32_t (&rb)[CL_MAXG], double* local, int ez, double v) {
        if (G > 1) {
#pragma unroll
            for (int r = 0; r < CL_MAXG; ++r)
                if (r < G) cl_st_f64(rb[r] + (uint32_t)ez * 8u, v);
        } else {
            local[ez] = v;
        }
    };
    // thread i < K holds the CTA's partial `mine` of value i: on return S.rv[i] = cluster total,
    // summed in rank order (deterministic, identical in every CTA).  Contains the barrier.
    auto reduce_post = [&](int K, double mine) {
        if (G == 1) {
            if (tid < K) S.rv[tid] = mine;
            __syncthreads();
            return;
        }
        if (tid < K) {
#pragma unroll
            for (int r = 0; r < CL_MAXG; ++r)
                if (r < G) cl_st_f64(rb_part[r] + (uint32_t)(((pbuf * G + rank) * CL_KRED + tid) * 8), mine);
        }
        cl_barrier();
        if (tid < K) {
            double acc = 0.0;
            for (int r = 0; r < G; ++r) acc += part[(pbuf * G + r) * CL_KRED + tid];
            S.rv[tid] = acc;
        }
        pbuf ^= 1;
        __syncthreads();
    };
    // block partial sums of K per-thread values, then the cluster reduction
    auto reduce_vals = [&](int K, const double* v) {
        double x[3];
        for (int k = 0; k < K; ++k) {
            x[k] = v[k];
            for (int o = 16; o > 0; o >>= 1) x[k] += __shfl_xor_sync(0xffffffffu, x[k], o);
        }
        if (lane == 0)
            for (int k = 0; k < K; ++k) S.red[warp * 4 + k] = x[k];
        __syncthreads();
        double mine = 0.0;
        if (tid < K)
            for (int w = 0; w < STB_NW; ++w) mine += S.red[w * 4 + tid];
        reduce_post(K, mine);
    };

    // ---- pair pass: F[slot][row] = mult * sum_k g(row, k) z[col_slot][k] for the own rows --------
    const int rr = lane & 7, kk = lane >> 3;
    // zb: offset of the gathered vector (Y.zy or Y.zv) in the dynamic shared array
    // lin_at >= 0: cooperative fields are linearised at the vector at offset lin_at (Krylov operator)
    auto fields_pass = [&](int zb, double mult, int lin_at) {
        const long long w_t0 = clock64();
        int g = my_ga, t0 = my_ta, piece = 0;
        while (g < NRG && (g < my_gb || (g == my_gb && t0 < my_tb))) {
            const int T = g >= nrg[0];
            const int gi = T ? g - nrg[0] : g;
            const int tend = g == my_gb ? my_tb : ST[T];
            const int i = min(gi * CL_RG + rr, cnt[T] - 1);
#pragma unroll
            for (int U = 0; U < 2; ++U) {
                const int lo = U ? stp[T][0] : 0;
                const int s0 = max(t0, lo) - lo, s1 = min(tend, lo + stp[T][U]) - lo;
                if (s0 >= s1) continue;
                const int ns = L.blk[T][U].n_slots;
                int dst, dpitch;                 // offsets into the dynamic shared array
                if (s0 == 0 && s1 == stp[T][U]) {
                    dst = Y.F + Y.foff[T] + (U ? L.blk[T][0].n_slots : 0) * Y.cntp[T] + gi * CL_RG + rr;
                    dpitch = Y.cntp[T];
                } else {
                    dst = Y.Pp + ((warp * 2 + (piece ? 1 : 0)) * Y.nsm) * CL_RG + rr;
                    dpitch = CL_RG;
                }
                if (gres_on[T][U])
                    cl_piece_ns<true>(ns, K.goff[T][U], nullptr, gpitch[T][U], i, zb + Y.zoff[U],
                                      Y.ldz[U], L.blk[T][U].slot_col, Y.nkp[U], kk, s0, s1, mult, dst, dpitch);
                else
                    cl_piece_ns<false>(ns, 0, gsrc[T][U], gpitch[T][U], i, zb + Y.zoff[U],
                                       Y.ldz[U], L.blk[T][U].slot_col, Y.nkp[U], kk, s0, s1, mult, dst, dpitch);
                ++piece;
            }
            ++g;
            t0 = 0;
        }
        if (lane == 0) S.wprof[warp] += clock64() - w_t0;
        CLPROF(0)
        __syncthreads();
        CLPROF(1)
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
            cl_coop_fields(L, Y, r_lo, cnt, zb, mult, lin_at, warp, lane);
            __syncthreads();
        }
        CLPROF(2)
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
    // x_a = sum_b Minv[a][b] r_b for own element (T, a, i), r = own-compact vector
    auto minv_apply = [&](int T, int a, int i, const double* r) -> double {
        const int s = L.s[T];
        if (quadpc && T == 0) {
            // row (i % pcg) * s + a of the inverse of the group's block; the blocks are stored TRANSPOSED
            // ([column][row][group]): the threads of a warp own different rows and read consecutive words
            const int pg = L.pcg, Q = i >> pg_sh, sq = Y.sq;
            const double* Bp = Qinv + ((i & (pg - 1)) * s + a) * Y.nqm + Q;
            const double* rq = r + ob[0] + pg * Q;
            double acc = 0.0;
            for (int m = 0; m < pg; ++m) {
                if (pg * Q + m >= cnt[0]) break;
                for (int b = 0; b < s; ++b) acc = fma(Bp[(m * s + b) * sq * Y.nqm], rq[b * cnt[0] + m], acc);
            }
            return acc;
        }
        const double* Mp = Minv + Y.moff[T] + (a * s) * Y.cntm[T] + i;
        const double* rq = r + ob[T] + i;
        double acc = 0.0;
        for (int b = 0; b < s; ++b) acc = fma(Mp[b * Y.cntm[T]], rq[b * cnt[T]], acc);
        return acc;
    };

    // f = rhs(z) for the own elements (odesolver.py:33-43, 65-71); fields stay in F
    auto rhs_pass = [&](const double* z, int zb, double* fout) {
        fields_pass(zb, 1.0, -1);
        for (int k = tid; k < n_own; k += STB_NT) {
            const int inf = own_inf[k];
            const int T = (inf >> 30) & 1, a = (inf >> 24) & 63, i = inf & 0xffffff;
            fout[k] = lin_sum(T, a, i, z) + term_sum(T, a, i, z);
        }
    };

    // block-Jacobi: Minv = (I - c (Lin_T + sum_t C_t F[slot_t] (e_b - e_a) e_a^T))^-1 per own ion.
    // 1. every thread builds its own row of M in place (thread per state, as the epilogues);
    // 2. the blocks are inverted in place: closed form for 2-state ions (thread per ion), else
    //    Gauss-Jordan with W = pow2 >= s lanes per ion, one row per lane.
    // refresh = false (lagged preconditioner): only the rows of the current M are rebuilt (into Mcur);
    // the inverses of the larger blocks stay those of an earlier step
    auto block_inverses = [&](double c, bool refresh) {
        // 1. rows of M = I - c J_qq: closed-form inverse for 1- and 2-state ions (thread per ion),
        //    else every thread builds the row of its own state in place
        for (int k = tid; k < n_own; k += STB_NT) {
            const int inf = own_inf[k];
            const int T = (inf >> 30) & 1, a = (inf >> 24) & 63, i = inf & 0xffffff;
            const int s = L.s[T], cm = Y.cntm[T];
            const TermTable& tt = L.terms[T];
            const double* lin = P.lin[T];
            const double* Fq = F + Y.foff[T] + i;
            if (s <= 2 && !(quadpc && T == 0)) {
                if (a != 0) continue;
                double m00 = 1.0 - c * lin[0], m01 = 0.0, m10 = 0.0, m11 = 1.0;
                if (s == 2) { m01 = -c * lin[1]; m10 = -c * lin[2]; m11 = 1.0 - c * lin[3]; }
                for (int t = 0; t < tt.n_terms; ++t) {
                    const double kq = c * P.C[tt.proc[t]] * Fq[tt.fslot[t] * Y.cntp[T]];
                    if (tt.a[t] == 0) { m00 += kq; m10 -= kq; }       // column a_t: row a_t +, row b_t -
                    else { m11 += kq; m01 -= kq; }
                }
                double* Mp = Minv + Y.moff[T] + i;
                if (s == 1) {
                    Mp[0] = m00 != 0.0 ? 1.0 / m00 : 0.0;
                } else {
                    const double det = m00 * m11 - m01 * m10;
                    const double id = det != 0.0 ? 1.0 / det : 0.0;
                    Mp[0] = m11 * id;
                    Mp[cm] = -m01 * id;
                    Mp[2 * cm] = -m10 * id;
                    Mp[3 * cm] = m00 * id;
                }
                continue;
            }
            double* Mp = (lagpc ? Mcur : Minv) + Y.moff[T] + (a * s) * cm + i;
            const double* lr = lin + a * s;
            for (int b = 0; b < s; ++b) Mp[b * cm] = (b == a ? 1.0 : 0.0) - c * lr[b];
            for (int n = S.tl_ptr[T][a]; n < S.tl_ptr[T][a + 1]; ++n) {
                const int e = S.tl_ent[T][n];
                const int t = e & 0x3fff;
                const double kq = c * P.C[tt.proc[t]] * Fq[tt.fslot[t] * Y.cntp[T]];
                Mp[tt.a[t] * cm] += (e & 0x4000) ? -kq : kq;     // row b_t: -, row a_t: +
            }
            if (lagpc && refresh && !(quadpc && T == 0)) {
                double* Mi = Minv + Y.moff[T] + (a * s) * cm + i;
                for (int b = 0; b < s; ++b) Mi[b * cm] = Mp[b * cm];
            }
        }
        __syncthreads();
        if (!refresh) return;
        if (quadpc && cnt[0] > 0) {
            // blocks of I - cJ over the quads: ion blocks from Mcur, couplings between the ions of a
            // quad from the pair rates, the state at the current iterate and the S-S processes:
            // interaction (donor q, partner k): rate = C g_qk y_q[ii] y_k[fi] moves q: ii -> if, k: fi -> ff
            const int s0 = L.s[0], sq = Y.sq, cm0 = Y.cntm[0], pg = L.pcg;
            const int nq = (cnt[0] + pg - 1) / pg;
            const double* g00 = L.blk[0][0].g[0];
            const double* yS = zy + Y.zoff[0] + r_lo[0];
            for (int e = tid; e < nq * sq * sq; e += STB_NT) {
                const int Q = e / (sq * sq), rc = e - Q * sq * sq;
                const int rho = rc / sq, gam = rc - rho * sq;
                const int mq = rho / s0, aq = rho - mq * s0, mk = gam / s0, ak = gam - mk * s0;
                const int iq = pg * Q + mq, ik = pg * Q + mk;
                double v;
                if (iq >= cnt[0] || ik >= cnt[0]) {
                    v = rho == gam ? 1.0 : 0.0;               // padding of the last quad
                } else if (mq == mk) {
                    v = Mcur[Y.moff[0] + (aq * s0 + ak) * cm0 + iq];
                } else {
                    v = 0.0;
                    const double g = g00 ? g00[(int64_t)(r_lo[0] + iq) * L.ld[0] + r_lo[0] + ik] : 0.0;
                    for (int p = 0; p < L.n_proc; ++p) {
                        const stb_process pr = A.procs[p];
                        if (pr.donor_type != 0 || pr.partner_type != 0 || P.C[p] == 0.0) continue;
                        const double cg = c * P.C[p] * g;
                        if (ak == pr.fi) {            // q donor, k partner: d rate / d y_k[fi]
                            const double w = cg * yS[pr.ii * Y.ldz[0] + iq];
                            if (aq == pr.ii) v += w;
                            if (aq == pr.if_) v -= w;
                        }
                        if (ak == pr.ii) {            // k donor, q partner: d rate / d y_k[ii]
                            const double w = cg * yS[pr.fi * Y.ldz[0] + iq];
                            if (aq == pr.fi) v += w;
                            if (aq == pr.ff) v -= w;
                        }
                    }
                }
                Qinv[(gam * sq + rho) * Y.nqm + Q] = v;       // transposed: (B^T)^-1 = (B^-1)^T
            }
            __syncthreads();
        }
        // 2. in-place Gauss-Jordan of the larger blocks, 64 threads per block (<= 3 entries each), no
        //    pivoting (column-diagonally-dominant M-matrix); compact rolled code: a lane-per-row
        //    shuffle version costs 4-10k cycles per call (tools/ubench.cu), mostly instruction fetch
        //    (blocks of more than 192 entries, the 32 x 32 group blocks: the whole CTA per block)
        auto gauss_jordan = [&](double* base, int s, int cm, int count) {
            const int ss = s * s, tpb = ss <= 192 ? 64 : STB_NT, tl = tid % tpb;
            int ei[3], ej[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int e = tl + tpb * u;
                ei[u] = e / s;
                ej[u] = e - ei[u] * s;
            }
            for (int blk0 = 0; blk0 < count; blk0 += STB_NT / tpb) {
                const int blk = blk0 + tid / tpb;
                const bool act = blk < count;
                double* Mq = base + (act ? blk : 0);
                for (int p = 0; p < s; ++p) {
                    double nv[3];
                    if (act) {
                        const double piv = Mq[(p * s + p) * cm];
                        const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            const int e = tl + tpb * u;
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
                            if (tl + tpb * u < ss) Mq[(tl + tpb * u) * cm] = nv[u];
                    }
                    __syncthreads();
                }
            }
        };
        for (int T = 0; T < 2; ++T) {
            if (L.s[T] <= 2 || cnt[T] <= 0 || (quadpc && T == 0)) continue;
            gauss_jordan(Minv + Y.moff[T], L.s[T], Y.cntm[T], cnt[T]);
        }
        if (quadpc && cnt[0] > 0) {
            const int nblk = (cnt[0] + L.pcg - 1) / L.pcg;
            if (Y.sq == 32) cl_gauss_jordan_warp32(Qinv, Y.nqm, nblk, warp, lane);
            else gauss_jordan(Qinv, Y.sq, Y.nqm, nblk);
        }
    };

    auto apply_change_D = [&](int order) {
        for (int k = tid; k < n_own; k += STB_NT) {
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
    const double inv_N = 1.0 / (double)N;
    const double tol_loose = A.lin_rms * sqrtN, tol_tight = fmin(A.lin_rms, STB_LIN_RMS) * sqrtN;
    const double rtol = A.rtol, atol = A.atol;

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
        if (S.cur_sys >= (A.sys_list ? A.n_list : A.n_sys)) break;
        const int sysl = A.sys_list ? A.sys_list[S.cur_sys] : S.cur_sys;
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
            C.newton_tol = fmax(10.0 * 2.220446049250313e-16 / rtol, fmin(0.03, sqrt(rtol)));
            C.tol_lin = tol_loose;
            C.lin_tight = 0;
            C.chk_jv = 0;
            C.chk_lin = 0;
            C.crate = 1.0;
            C.since_verify = 0;
            C.gfac = 0.0;
            C.pc_c = 0.0;
            C.pc_age = -1;
            C.pc_refresh = 1;
            memset(&C.st, 0, sizeof(stb_stats));
        }
        int64_t k_out_l = 1;                                   // output cursor, uniform over the cluster
        double tq_l = A.n_t > 1 ? A.t[1] : INFINITY;
        double tq_n = A.n_t > 2 ? A.t[2] : INFINITY;           // one output time ahead: its load latency stays off the step
        // D0 = y0, row 0 of the output = y0 verbatim (odesolver.py:107)
        for (int k = tid; k < n_own; k += STB_NT) {
            const int inf = own_inf[k];
            const int T = (inf >> 30) & 1, a = (inf >> 24) & 63, i = inf & 0xffffff;
            const int e = L.off[T] + a * L.n[T] + r_lo[T] + i;
            const double v = y0[e];
            D[k] = v;
            for (int j = 1; j < 8; ++j) D[j * PO + k] = 0.0;
            y_out[own_ref[k]] = v;
            push(rb_zy, zy, own_ez[k], v);
        }
        sync_all();
        rhs_pass(zy, Y.zy, fv);
        // ---- initial step (Hairer-Wanner, as scipy select_initial_step) -----------------------
        {
            double v3[3] = {0.0, 0.0, 0.0};
            for (int k = tid; k < n_own; k += STB_NT) {
                const double y = zy[own_ez[k]];
                const double sc = atol + rtol * fabs(y);
                scl[k] = sc;
                const double a = y / sc, b = fv[k] / sc;
                v3[0] = fma(a, a, v3[0]);
                v3[1] = fma(b, b, v3[1]);
                if (!isfinite(fv[k])) v3[2] = 1.0;
            }
            reduce_vals(3, v3);
            const double d0 = sqrt(S.rv[0] / N), d1 = sqrt(S.rv[1] / N);
            const double bad = S.rv[2];
            const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            __syncthreads();
            for (int k = tid; k < n_own; k += STB_NT)
                push(rb_zv, zv, own_ez[k], zy[own_ez[k]] + h0 * fv[k]);
            sync_all();
            rhs_pass(zv, Y.zv, tw);
            double v1[1] = {0.0};
            for (int k = tid; k < n_own; k += STB_NT) {
                const double a = (tw[k] - fv[k]) / scl[k];
                v1[0] = fma(a, a, v1[0]);
            }
            reduce_vals(1, v1);
            if (tid == 0) {
                C.st.n_rhs += 2;
                const double d2 = sqrt(S.rv[0] / N) / h0;
                const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3)
                                                               : sqrt(0.01 / fmax(d1, d2));
                C.h = fmin(fmin(100.0 * h0, h1), C.t_end - C.t);
                if (bad != 0.0 || !isfinite(C.h) || !(C.h > 0.0)) C.status = 3;
            }
            __syncthreads();
            const double h = C.h;
            for (int k = tid; k < n_own; k += STB_NT) D[PO + k] = fv[k] * h;
        }
        __syncthreads();

        // =============================== time stepping =====================================
        while (true) {
            __syncthreads();
            CLPROF(14)
            if (C.status != 0 || k_out_l >= A.n_t) break;
            // scalar prologue, evaluated by every thread from the same shared values (no serial section,
            // no second barrier on the common path)
            double h = C.h;
            const double t_cur = C.t, t_end = C.t_end;
            const int order = C.order;
            {
                const double t_try = t_cur + h;
                if (t_try > t_end && (t_try - t_end) > 1e-14 * fabs(t_end)) {
                    // the last step is shortened to land on t_end (rare)
                    __syncthreads();
                    if (tid == 0) {
                        const double factor = (t_end - t_cur) / h;
                        cl_compute_RU(order, factor, C.RU, K.Rs, K.Us);
                        C.h = t_end - t_cur;
                        C.n_equal = 0;
                        C.gfac *= cl_ipow(factor, order + 1);
                    }
                    __syncthreads();
                    apply_change_D(order);
                    h = t_end - t_cur;
                } else if (t_try > t_end) {
                    h = t_end - t_cur;      // round-off sized overshoot: land exactly on t_end
                }
            }
            {
                int st_now = 0;
                if (h < 10.0 * 2.220446049250313e-16 * fabs(t_cur)) st_now = 2;
                if (C.steps_since_out >= A.nsteps) st_now = 1;
                if (st_now) {
                    __syncthreads();
                    if (tid == 0) C.status = st_now;
                    break;
                }
            }
            const double tol_lin = C.tol_lin;      // loose, or tight for a while after a rejection
            const double ralpha = K.RALPHA[order];
            const double c = h * ralpha;
            const bool pc_ref = !lagpc || C.pc_age < 0 || C.pc_age >= A.pc_maxage ||
                                fabs(c - C.pc_c) > A.pc_drift * fabs(C.pc_c);
            const double gf = guess ? C.gfac : 0.0;
            bool first_newton = true;
            // predictor, scale, psi; d = 0; the predicted state goes to every CTA of the cluster
            for (int k = tid; k < n_own; k += STB_NT) {
                double yp = 0.0, ps = 0.0;
                for (int j = 0; j <= order; ++j) {
                    const double dj = D[j * PO + k];
                    yp += dj;
                    if (j >= 1) ps = fma(GAMMA[j], dj, ps);
                }
                scl[k] = atol + rtol * fabs(yp);
                psi[k] = ps * ralpha;
                // Newton start: predictor + the previous correction (D[order + 1]) rescaled to this step
                const double d0 = gf != 0.0 ? gf * D[(order + 1) * PO + k] : 0.0;
                dd[k] = d0;
                push(rb_zy, zy, own_ez[k], yp + d0);
            }
            if (tid == 0) { C.newton_it = 0; C.flag = 0; C.dy_norm_old = -1.0; }
            // ---- Newton-Krylov ------------------------------------------------------------------
            while (true) {       // flag: 0 iterate, 1 converged, 2 failed
                CLPROF(15)
                sync_all();      // y_k complete in every CTA
                CLPROF(3)
                rhs_pass(zy, Y.zy, fv);
                __syncthreads();
                CLPROF(4)
                const bool pc_now = !lagpc || (pc_ref && first_newton);
                block_inverses(c, pc_now);
                __syncthreads();
                CLPROF(5)
                // tw = Minv (c f - psi - d) / scale; the unscaled vector goes to zv
                {
                    for (int k = tid; k < n_own; k += STB_NT) kv[k] = c * fv[k] - psi[k] - dd[k];
                    __syncthreads();
                    double v2[2] = {0.0, 0.0};
                    for (int k = tid; k < n_own; k += STB_NT) {
                        const int inf = own_inf[k];
                        const int T = (inf >> 30) & 1, a = (inf >> 24) & 63, i = inf & 0xffffff;
                        const double xy = minv_apply(T, a, i, kv);
                        const double x = xy / scl[k];
                        tw[k] = x;
                        push(rb_zv, zv, own_ez[k], xy);
                        v2[0] = fma(x, x, v2[0]);
                        if (!isfinite(fv[k])) v2[1] = 1.0;
                    }
                    CLPROF(6)
                    reduce_vals(2, v2);
                    CLPROF(7)
                }
                // beta and the start of the linear solve are evaluated by every thread from the reduced
                // values; only the rare closures (non-finite residual, residual already below the linear
                // tolerance) go through thread 0 and a barrier
                const double beta0 = sqrt(S.rv[0]);
                const bool bad0 = S.rv[1] != 0.0 || !isfinite(beta0);
                const bool small0 = !bad0 && !(beta0 > tol_lin);
                if (tid == 0) {
                    if (lagpc && pc_now) { C.pc_c = c; C.pc_age = 0; }
                    C.st.n_rhs += 1;
                    C.st.n_lin += 1;
                    C.beta = beta0;
                    C.kk = 0;
                    S.gg[0] = beta0;
                    if (bad0) C.flag = 2;
                    if (small0) {
                        // residual already below the linear tolerance: dy = 0
                        int fl = newton_decide(C, 0.0);
                        if (fl == 0 && C.newton_it >= STB_NEWTON_MAXIT) fl = 2;
                        C.flag = fl;
                    }
                }
                if (bad0 || small0) __syncthreads();
                // ---- GMRES: V_0 = tw / beta; zv holds the direction, scaled by `zmul` ----------
                // (C.flag: 0 while the solve runs; the thread that closes the solve leaves the Newton
                // decision there: 1 converged, 2 failed, 0 iterate again)
                bool lin_done = bad0 || small0;
                int j = 0;
                double zmul = beta0 > 0.0 ? 1.0 / beta0 : 0.0;
                if (!lin_done) {
                    double* V0 = Vrow(0);
                    for (int k = tid; k < n_own; k += STB_NT) V0[k] = tw[k] * zmul;
                }
                while (!lin_done) {
                    // w = Minv (I - cJ) v / scale = V_j - c Minv k(v) / scale, k = inter-ion coupling
                    CLPROF(15)
                    fields_pass(Y.zv, zmul, Y.zy);
                    for (int k = tid; k < n_own; k += STB_NT) {
                        const int inf = own_inf[k];
                        const int T = (inf >> 30) & 1, a = (inf >> 24) & 63, i = inf & 0xffffff;
                        double kq = term_sum(T, a, i, zy);
                        if (lagpc && (L.s[T] > 2 || (quadpc && T == 0))) {
                            // lagged inverse: Minv_old (M v - c k(v)), rows of the current M in Mcur
                            const int s = L.s[T], cm = Y.cntm[T];
                            const double* Mp = Mcur + Y.moff[T] + (a * s) * cm + i;
                            const double* vq = zv + Y.zoff[T] + r_lo[T] + i;
                            double mv = 0.0;
                            for (int b = 0; b < s; ++b) mv = fma(Mp[b * cm], vq[b * Y.ldz[T]], mv);
                            kq = mv * zmul - c * kq;
                        }
                        kv[k] = kq;
                    }
                    __syncthreads();
                    {
                        const double* Vj = Vrow(j);
                        for (int k = tid; k < n_own; k += STB_NT) {
                            const int inf = own_inf[k];
                            const int T = (inf >> 30) & 1, a = (inf >> 24) & 63, i = inf & 0xffffff;
                            const double mk = minv_apply(T, a, i, kv);
                            tw[k] = lagpc && (L.s[T] > 2 || (quadpc && T == 0)) ? mk / scl[k]
                                                                                : Vj[k] - c * mk / scl[k];
                        }
                    }
                    __syncthreads();
                    CLPROF(8)
                    // h_i = <w, V_i> (i <= j), ww = <w, w>: warp i takes dot i
                    const int K = j + 2;
                    auto dots = [&]() {
                        for (int i = warp; i < K; i += STB_NW) {
                            const double* Vi = i <= j ? Vrow(i) : tw;
                            double acc = 0.0;
                            for (int k = lane; k < n_own; k += 32) acc = fma(tw[k], Vi[k], acc);
                            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                            if (lane == 0) S.rv[i] = acc;
                        }
                        __syncthreads();
                    };
                    auto orth = [&]() {
                        for (int k = tid; k < n_own; k += STB_NT) {
                            double acc = tw[k];
                            for (int i = 0; i <= j; ++i) acc = fma(-S.rv[i], Vrow(i)[k], acc);
                            tw[k] = acc;
                        }
                    };
                    dots();
                    CLPROF(9)
                    reduce_post(K, tid < K ? S.rv[tid] : 0.0);
                    CLPROF(10)
                    const double ww = S.rv[j + 1];
                    double hn2 = ww;
                    for (int i = 0; i <= j; ++i) hn2 = fma(-S.rv[i], S.rv[i], hn2);
                    orth();
                    // Pythagorean norm of the orthogonalised vector: with ||w_orth|| > 1e-4 ||w|| the
                    // cancellation costs at most 8 of 16 digits of hn2 and the basis stays orthogonal
                    // to ~1e-12, far inside the linear tolerance; below that (a preconditioned operator
                    // this close to the identity ends the solve anyway) re-orthogonalise explicitly
                    const bool reorth = !(hn2 > 1e-8 * ww);
                    if (reorth) {
                        // cancellation: explicit norm and one re-orthogonalisation pass (DGKS)
                        if (tid <= j) S.H[tid * CL_KRYLOV + j] = S.rv[tid];
                        __syncthreads();
                        dots();
                        reduce_post(K, tid < K ? S.rv[tid] : 0.0);
                        const double w2 = S.rv[j + 1];
                        hn2 = w2;
                        for (int i = 0; i <= j; ++i) hn2 = fma(-S.rv[i], S.rv[i], hn2);
                        if (tid <= j) S.H[tid * CL_KRYLOV + j] += S.rv[tid];
                        orth();
                        if (!(hn2 > 1e-8 * w2)) {
                            double v1[1] = {0.0};
                            for (int k = tid; k < n_own; k += STB_NT) v1[0] = fma(tw[k], tw[k], v1[0]);
                            reduce_vals(1, v1);
                            hn2 = S.rv[0];
                        }
                        __syncthreads();     // H column complete
                    }
                    const double ihn = hn2 > 0.0 ? rsqrt(hn2) : 0.0;
                    const double hn = hn2 * ihn;
                    // Givens update of column j, residual estimate and (when the solve closes) the
                    // triangular solve + Newton decision: one thread of the least loaded warp, off
                    // the path of the threads that orthogonalise and exchange the next direction
                    if (tid == STB_NT - 32) {
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
                                for (int k = i + 1; k < nk; ++k) acc -= S.H[i * CL_KRYLOV + k] * S.yk_s[k];
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
                    {
                        double* Vn = Vrow(j + 1);
                        for (int k = tid; k < n_own; k += STB_NT) {
                            const double v = tw[k] * ihn;
                            Vn[k] = v;
                            push(rb_zv, zv, own_ez[k], v * scl[k]);
                        }
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
                if (flag != 2) {
                    for (int k = tid; k < n_own; k += STB_NT) {
                        double x = 0.0;
                        for (int i = 0; i < nk; ++i) x = fma(S.yk_s[i], Vrow(i)[k], x);
                        const double dy = x * scl[k];
                        const double yn = zy[own_ez[k]] + dy;
                        dd[k] += dy;
                        if (flag == 0) push(rb_zy, zy, own_ez[k], yn);   // next iteration needs it
                        else zy[own_ez[k]] = yn;
                    }
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
                    C.lin_tight = A.lin_tight_steps;
                    C.tol_lin = tol_tight;
                }
                __syncthreads();
                apply_change_D(order);
                continue;
            }
            CLPROF(13)
            const double t_new = t_cur + h;
            // ---- error test and order selection share one reduction -------------------------------
            {
                double v3[3] = {0.0, 0.0, 0.0};
                for (int k = tid; k < n_own; k += STB_NT) {
                    const double sc = atol + rtol * fabs(zy[own_ez[k]]);
                    scl[k] = sc;
                    const double rsc = 1.0 / sc;
                    const double dv = dd[k];
                    const double er = ERRC[order] * dv * rsc;
                    v3[0] = fma(er, er, v3[0]);
                    if (order > 1) {
                        const double em = ERRC[order - 1] * (D[order * PO + k] + dv) * rsc;
                        v3[1] = fma(em, em, v3[1]);
                    }
                    if (order < STB_MAXORD) {
                        const double ep = ERRC[order + 1] * (dv - D[(order + 1) * PO + k]) * rsc;
                        v3[2] = fma(ep, ep, v3[2]);
                    }
                }
                reduce_vals(3, v3);
            }
            // accept / reject evaluated by every thread from the reduced values (no serial section)
            const double err_norm = sqrt(S.rv[0] * inv_N);
            const double err_m2 = S.rv[1], err_p2 = S.rv[2];
            const double safety = 0.9 * (2 * STB_NEWTON_MAXIT + 1) / (2.0 * STB_NEWTON_MAXIT + C.n_iter);
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
                    C.lin_tight = A.lin_tight_steps;
                    C.tol_lin = tol_tight;
                }
                __syncthreads();
                apply_change_D(order);
                continue;
            }
            // ---- accepted: update differences, dense output, order / step selection -------------
            // Every thread updates and interpolates its own elements with no barrier in between
            // (the output cursor k_out_l / tq_l is kept identically by all threads); the scalar
            // control work runs meanwhile on the last, least loaded warp.
            for (int k = tid; k < n_own; k += STB_NT) {
                const double dv = dd[k];
                const double dp1 = D[(order + 1) * PO + k];
                D[(order + 2) * PO + k] = dv - dp1;
                D[(order + 1) * PO + k] = dv;
                double run = dv;
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
                    double pj[STB_MAXORD];
                    double pr = 1.0;
#pragma unroll
                    for (int jj = 0; jj < STB_MAXORD; ++jj) {
                        pr *= (tq_l - (t_new - h * jj)) * (ih * RCPJ[jj]);
                        pj[jj] = jj < order ? pr : 0.0;
                    }
                    double* row = y_out + k_out_l * (int64_t)N;
                    for (int k = tid; k < n_own; k += STB_NT) {
                        double acc = D[k];
                        if (!exact) {
#pragma unroll
                            for (int jj = 0; jj < STB_MAXORD; ++jj)
                                if (jj < order) acc = fma(pj[jj], D[(jj + 1) * PO + k], acc);
                        }
                        row[own_ref[k]] = acc;
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
                        const double en = lane == 1 ? err_norm : sqrt((lane == 0 ? err_m2 : err_p2) * inv_N);
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
                    if (A.restart && C.st.n_steps % A.restart_steps == 0 && C.st.n_steps <= CL_RESTART_LAST) {
                        // iterations per linear solve over the last restart_steps steps
                        if ((double)(C.st.n_jv - C.chk_jv) > A.restart_its * (double)(C.st.n_lin - C.chk_lin)) {
                            C.status = 4;
                            *A.restart = 1;
                        }
                        C.chk_jv = C.st.n_jv;
                        C.chk_lin = C.st.n_lin;
                    }
                    C.steps_since_out = wrote ? 1 : C.steps_since_out + 1;
                    C.k_out = k_out_l;
                    C.st.n_out = k_out_l;
                    C.t = t_new;
                    C.flag2 = 0;      // (not C.flag: slow threads may still be reading the rejection test)
                    if (C.pc_age >= 0) C.pc_age += 1;
                    if (C.lin_tight > 0 && --C.lin_tight == 0) C.tol_lin = tol_loose;
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
            for (int k = 0; k < 24; ++k) A.prof[k] = S.prof[k];
            for (int k = 0; k < STB_NW; ++k) A.prof[24 + k] = S.wprof[k];
            for (int k = 0; k <= STB_NW; ++k) A.prof[24 + STB_NW + k] = S.wpos[k];
        }
        if (tid == 0 && rank == 0) {
            C.st.status = C.status;
            C.st.t_reached = C.t;
            C.st.n_out = C.k_out;
            A.stats[sysl] = C.st;
        }
    }
}
