// Persistent device-resident stiff integrator: variable-order (1-5) BDF in the
// quasi-constant-step NDF form, Jacobian-free Newton-Krylov (left-preconditioned, scaled
// GMRES(30)) with a per-ion block-Jacobi preconditioner.  Replaces odesolver._solve_ode
// (odesolver.py:85-137) and the scipy VODE/LSODA call behind it, with no host round trip per
// step: the grid is split into groups of G CTAs, each group integrates whole ODE systems of the
// batch one after the other; CTAs of a group own fixed sets of ion rows, exchange only the
// gathered state vector, and synchronise through a monotonic counter in global memory
// (co-residency is guaranteed by the cooperative launch).  Every CTA runs the (tiny) scalar
// step/order/Newton/Givens logic redundantly from identical reduced values, so control flow is
// uniform across the group by construction.
#pragma once
#include "stb_coop.cuh"
#include "stb_tma.cuh"

#define STB_KRED (STB_KRYLOV + 4)
#define STB_NEWTON_MAXIT 4
#define STB_MIN_FACTOR 0.2
#define STB_MAX_FACTOR 10.0
// Linear (GMRES) tolerance: preconditioned residual, RMS over the states in units of the error
// test (1 = the local error bound).  The Newton update of a step is accepted with this residual,
// so it plays the role of LSODA's corrector tolerance (conit ~ 0.07 of the error bound).
#define STB_LIN_RMS 2e-3
// The cluster integrator runs with a looser one: with the quad block preconditioner an iteration is
// worth more than a few rejected steps.  Measured on the C2 sweep (profiles/r02_lin_tol.txt): 673 solves/s
// at 2e-3, 739 at 5e-2 ... 1e-1 (LSODA's corrector tolerance is 0.07-0.17 of the error bound), 634 at
// 0.2 (rejections 24 -> 787); parity metric of the states 0.03 -> 0.07 of the allowed 1e-4 |ref| + 1e-12.
#define CL_LIN_RMS 5e-2
// ... but not blindly: on stiffer systems (the reference's 240S/108A benchmark lattice) sloppy
// solves feed the error test and the step count explodes (2 102 -> 16 693 steps at 5e-2).  After a
// rejected step or a failed Newton iteration the next CL_LIN_TIGHT_STEPS accepted steps use STB_LIN_RMS
// (240S/108A: 489 -> 511 ms; C2 sweep 675 -> 731 solves/s; profiles/r02_lin_tol.txt).
#define CL_LIN_TIGHT_STEPS 40
#define GRID_LIN_RMS 2e-2      // grid integrator (with the same tightening rule): -6..7 % on the C1 lattices, flat
                               // between 1e-2 and 5e-2, parity metric unchanged (profiles/r02_lin_tol.txt)
// quads -> 16-ion groups (see stb_system_create): iterations per linear solve over windows of this many steps
#define CL_RESTART_STEPS 48
#define CL_RESTART_ITS 10.0
#define CL_BATCH_ITS 8.0
#define CL_RESTART_LAST 480
#define STB_RESTART_GROUPS 1000     // internal return code of integrate_cluster
// The Newton convergence rate is carried across steps (LSODA / CVODE `crate`): a step whose first
// update already satisfies 1.5 * min(1, crate) * ||dy|| < newton_tol is accepted without a second
// residual evaluation; every STB_VERIFY_EVERY such steps a second iteration is forced, which
// re-measures the rate from the true nonlinear residual.
#define STB_VERIFY_EVERY 16

struct IntArgs {
    const DevLattice* L;
    const double* blocks;
    const double* strengths;
    int first, n_sys, pulse;
    int G, n_groups;
    int64_t n_t;
    const double* t;          // [n_t]
    const double* y0;         // [n_sys][N] planes
    double rtol, atol;
    int64_t nsteps;
    double* y_out;            // [n_sys][n_t][N] reference layout
    stb_stats* stats;         // [n_sys]
    double* work;             // [n_groups][work_stride]
    int64_t work_stride;
    unsigned long long* bar;  // [n_groups]
    double* partials;         // [n_groups][2][G][STB_KRED]
    int NMinv;                // doubles of block-Jacobi inverses per system
    int resident;             // the CTA's pair rows are copied to shared memory once
    int staged;               // gather vector staged in shared memory (N*8 bytes dynamic smem)
    long long* prof;          // [n_sys][8] clock64 cycles per phase (thread 0 of CTA 0), may be null
    int tma;                  // dense rows fed by bulk copies through a shared-memory ring (stb_tma.cuh)
    int tma_kc, tma_stages;
    int guess;                // Newton starts from the previous step's correction (see CL_OPT_GUESS)
    double lin_rms;           // GMRES tolerance in units of the error test (GRID_LIN_RMS), tightened to
    int lin_tight_steps;      // STB_LIN_RMS for this many steps after a rejection (as the cluster integrator)
    // group block preconditioner: the sensitizers are stored in compact groups of `grp` ions
    // (grp * s_S = STB_GDIM unknowns) and the block of I - cJ over a whole group is inverted instead
    // of the single-ion blocks (0: off).  NGinv doubles per system hold the inverses.
    int grp, NGinv;
    const stb_process* procs;
    int lag;                  // keep the inverses of the group blocks across steps (STB_PC_MAXAGE, STB_PC_DRIFT)
    int a_ctas;               // TMA passes: activator rows on CTAs of their own
};
#define STB_PC_MAXAGE 20
#define STB_PC_DRIFT 0.3
#define STB_GDIM 32           // unknowns of a group block: one warp applies it, lane = row

__device__ __forceinline__ double stb_ipow(double x, int n) {
    double r = 1.0;
    for (int i = 0; i < n; ++i) r *= x;
    return r;
}

__host__ __device__ inline int64_t stb_work_doubles(int N, int NF, int NMinv) {
    // D[8] yk d psi scale f wv gv wbuf  V[KRYLOV+1]  F  Minv
    return (int64_t)N * (8 + 8 + STB_KRYLOV + 1) + NF + NMinv;
}

struct Ctl {
    double t, h, t_end, c, newton_tol, tol_lin;
    double beta, inrm, dy_norm, dy_norm_old, err_norm, safety, crate;
    double gfac;              // scale of the previous step's correction used as the Newton start
    double pc_c;              // c of the step the lagged block inverses were factored at
    int pc_age, pc_refresh;   // steps since then; factor them in this Newton iteration
    int lin_tight;            // accepted steps still to be solved with the tight GMRES tolerance
    long long chk_jv, chk_lin;   // counters at the last check of the restart rule (CL_RESTART_STEPS)
    double RU[(STB_MAXORD + 1) * (STB_MAXORD + 1)];
    double p[STB_MAXORD + 1];
    int order, n_equal, status, kk, flag, flag2, newton_it, n_iter, since_verify;
    int64_t k_out, steps_since_out;
    stb_stats st;
    long long prof[8];
};

// Newton control after an update of scaled RMS norm dy_norm (thread 0).  Returns 0: iterate,
// 1: converged, 2: failed (diverging).  The tests are scipy BDF's (solve_bdf_system) plus the
// carried-rate acceptance described at STB_VERIFY_EVERY.
__device__ __forceinline__ int newton_decide(Ctl& C, double dy_norm) {
    const int it = C.newton_it;
    const double rate = C.dy_norm_old > 0.0 ? dy_norm / C.dy_norm_old : -1.0;
    if (rate >= 0.0 && isfinite(rate)) {
        C.crate = fmax(0.2 * C.crate, rate);
        C.since_verify = 0;
    }
    int flag = 0;
    if (!isfinite(dy_norm)) flag = 2;
    else if (rate >= 0.0 &&
             (rate >= 1.0 ||
              pow(rate, (double)(STB_NEWTON_MAXIT - it)) / (1.0 - rate) * dy_norm > C.newton_tol))
        flag = 2;
    C.dy_norm = dy_norm;
    if (flag == 0) {
        if (dy_norm == 0.0 || (rate >= 0.0 && rate / (1.0 - rate) * dy_norm < C.newton_tol) ||
            (rate < 0.0 && dy_norm < 0.1 * C.newton_tol)) {
            flag = 1;
        } else if (rate < 0.0 && C.since_verify < STB_VERIFY_EVERY &&
                   1.5 * fmin(1.0, C.crate) * dy_norm < C.newton_tol) {
            flag = 1;
            C.since_verify += 1;
        }
    }
    C.dy_norm_old = dy_norm;
    C.newton_it = it + 1;
    C.n_iter = it + 1;
    return flag;
}

struct GroupCtx {
    unsigned long long* counter;
    unsigned long long epoch;
    double* partials;   // [2][G][STB_KRED]
    int G, cta, buf, relaxed;
};

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

// Group barrier.  Two flavours:
//  * acquire: fence + acquire load, so that plain (L1-cached) loads issued afterwards observe
//    the other CTAs' writes (the fence invalidates this SM's L1);
//  * relaxed: release on arrival only.  Data written by other CTAs is then read exclusively
//    through L2 (`__ldcg`, staging copies), so the L1 - which holds this CTA's own rows - stays
//    warm across barriers.  Used when the gathered vector is staged in shared memory.
__device__ __forceinline__ void group_sync(GroupCtx& g) {
    __syncthreads();
    if (g.G > 1) {
        if (threadIdx.x == 0) {
            g.epoch += 1;
            const unsigned long long target = g.epoch * (unsigned long long)g.G;
            if (g.relaxed) {
                asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(g.counter), "l"(1ull)
                             : "memory");
                while (ld_relaxed_u64(g.counter) < target) { }
            } else {
                __threadfence();
                atomicAdd(g.counter, 1ull);
                while (ld_acquire_u64(g.counter) < target) { }
                __threadfence();
            }
        }
        __syncthreads();
    }
}

// vals[0..K) hold this CTA's partial sums (smem); on return they hold the group totals,
// summed in CTA order (deterministic, identical in every CTA).
__device__ __forceinline__ void group_reduce(GroupCtx& g, double* vals, int K) {
    if (g.G == 1) { __syncthreads(); return; }
    double* mine = g.partials + ((int64_t)g.buf * g.G + g.cta) * STB_KRED;
    if ((int)threadIdx.x < K) __stcg(mine + threadIdx.x, vals[threadIdx.x]);
    group_sync(g);
    if ((int)threadIdx.x < K) {
        const double* base = g.partials + (int64_t)g.buf * g.G * STB_KRED + threadIdx.x;
        double acc = 0.0;
        for (int c = 0; c < g.G; ++c) acc += __ldcg(base + (int64_t)c * STB_KRED);
        vals[threadIdx.x] = acc;
    }
    g.buf ^= 1;
    __syncthreads();
}

template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* red /*[NW*K]*/, double* outs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k)
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[warp * K + k] = v[k];
    }
    __syncthreads();
    if ((int)threadIdx.x < K) {
        double acc = 0.0;
        for (int w = 0; w < STB_NW; ++w) acc += red[w * K + threadIdx.x];
        outs[threadIdx.x] = acc;
    }
    __syncthreads();
}

__device__ __forceinline__ void compute_R(int order, double factor, double* R /*6x6*/) {
    const int n = order + 1;
    for (int j = 0; j < n; ++j) R[j] = 1.0;
    for (int i = 1; i < n; ++i) {
        R[i * 6] = 0.0;
        for (int j = 1; j < n; ++j)
            R[i * 6 + j] = R[(i - 1) * 6 + j] * ((i - 1 - factor * j) / i);
    }
}

// RU = R(order, factor) . R(order, 1)   (thread 0)
__device__ __forceinline__ void compute_RU(int order, double factor, double* RU) {
    double R[36], U[36];
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

#define OWN_LOOP(e)                                                          \
    for (int T_ = 0; T_ < 2; ++T_) {                                         \
        const int cnt_ = r_hi[T_] - r_lo[T_];                                \
        const int tot_ = cnt_ * L.s[T_];                                     \
        for (int k_ = threadIdx.x; k_ < tot_; k_ += STB_NT) {                \
            const int e = L.off[T_] + (k_ / cnt_) * L.n[T_] + r_lo[T_] + (k_ % cnt_);
#define OWN_END }}

__global__ void __launch_bounds__(STB_NT, 1) k_integrate(IntArgs A) {
    __shared__ DevLattice L;
    __shared__ SysParams P;
    __shared__ Ctl C;
    __shared__ double H[(STB_KRYLOV + 1) * STB_KRYLOV];
    __shared__ double cs[STB_KRYLOV], sn[STB_KRYLOV], gg[STB_KRYLOV + 1], yk_s[STB_KRYLOV];
    __shared__ double red[STB_NW * 4];
    __shared__ double rv[STB_KRED];
    __shared__ double Gs[STB_GDIM * (STB_GDIM + 1)];    // group block being assembled / inverted
    __shared__ int ss_n, ss_st[STB_MAX_PROC][4];        // active S-S processes of the system (ii, fi, if, ff) ...
    __shared__ double ss_C[STB_MAX_PROC];               // ... and their strengths: the couplings inside a group
    STB_DYN_SMEM(zs);                // staged gather vector (N doubles) when A.staged, or the TMA ring

    const int gid = blockIdx.x / A.G;
    if (gid >= A.n_groups) return;
    for (int i = threadIdx.x; i < (int)(sizeof(DevLattice) / 4); i += blockDim.x)
        ((int*)&L)[i] = ((const int*)A.L)[i];
    __syncthreads();

    GroupCtx g;
    g.G = A.G;
    g.cta = blockIdx.x % A.G;
    g.counter = A.bar + gid;
    g.epoch = 0;
    g.buf = 0;
    g.relaxed = A.staged;
    g.partials = A.partials + (int64_t)gid * 2 * A.G * STB_KRED;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = L.N;
    int sl_lo[2], sl_hi[2], r_lo[2], r_hi[2];
    for (int T = 0; T < 2; ++T) {
        sl_lo[T] = (int)(((int64_t)L.n_slices[T] * g.cta) / g.G);
        sl_hi[T] = (int)(((int64_t)L.n_slices[T] * (g.cta + 1)) / g.G);
        r_lo[T] = min(sl_lo[T] * L.R, L.n[T]);
        r_hi[T] = min(sl_hi[T] * L.R, L.n[T]);
    }
    // TMA passes: the few activator rows go to CTAs of their own.  Dealt over all CTAs (35 activators on 148
    // CTAs at N_uc 20) they give a quarter of the CTAs one more row tile - a tile of one row costs a whole sweep
    // over the partner columns - and every pass waits for those: the arrival skew at the group barriers was one
    // tile time.  So, while a CTA has at most four tiles of sensitizer rows, nA CTAs take one tile of activator
    // rows each and no sensitizer row (N_uc 20: 207 -> 197 ms).  Beyond that the extra one-row tile is a small
    // part of a pass and an activator row costs several sensitizer rows (more field slots): dedicated CTAs
    // become the slowest ones (N_uc 30 with one CTA for all 73 activators: 624 -> 938 ms), so the rows stay dealt.
    int cS = g.cta, GS = g.G;            // this CTA's index among the CTAs that own sensitizer rows, their number
    const int nA_want = (L.n[1] + STB_TMA_ROWS - 1) / STB_TMA_ROWS;
    if (A.tma && A.a_ctas && L.n[0] > 0 && L.n[1] > 0 && 2 * nA_want <= g.G &&
        (L.n[0] + g.G * STB_TMA_ROWS - 1) / (g.G * STB_TMA_ROWS) <= 4) {
        const int nA = nA_want;
        GS = g.G - nA;
        if (g.cta < nA) {
            cS = -1;
            r_lo[1] = (int)(((int64_t)L.n[1] * g.cta) / nA);
            r_hi[1] = (int)(((int64_t)L.n[1] * (g.cta + 1)) / nA);
            r_lo[0] = r_hi[0] = 0;
        } else {
            cS = g.cta - nA;
            r_lo[0] = (int)(((int64_t)L.n[0] * cS) / GS);
            r_hi[0] = (int)(((int64_t)L.n[0] * (cS + 1)) / GS);
            r_lo[1] = r_hi[1] = L.n[1];
        }
        for (int T = 0; T < 2; ++T) {       // (TMA passes: one row per slice)
            sl_lo[T] = r_lo[T] / L.R;
            sl_hi[T] = (r_hi[T] + L.R - 1) / L.R;
        }
    }
    // group preconditioner: a CTA owns - as vectors: Krylov basis, dot products, preconditioner - whole
    // groups of sensitizers (type-local rows grp*m .. grp*m + grp-1, see group_sensitizers: what counts is
    // that no strongly coupled pair is cut by a group boundary).  The PAIR PASS keeps the split balanced to
    // a row (pr_lo .. pr_hi: whole groups per CTA cost 14-20 % of a pass on the C1 lattices): rows near a
    // boundary are streamed by one CTA and post-processed by its neighbour, which waits for it
    // (neighbour_sync) - both ranges differ by less than a group, so only adjacent CTAs are involved.
    const int grp = A.grp;
    int n_ch = 0, ch_slot = 0;
    int pr_lo[2] = {r_lo[0], r_lo[1]}, pr_hi[2] = {r_hi[0], r_hi[1]};
    bool split = false;
    if (grp && cS >= 0) {
        const int ng = (L.n[0] + grp - 1) / grp;
        ch_slot = (int)(((int64_t)ng * cS) / GS);
        n_ch = (int)(((int64_t)ng * (cS + 1)) / GS) - ch_slot;
        r_lo[0] = min(ch_slot * grp, L.n[0]);
        r_hi[0] = min((ch_slot + n_ch) * grp, L.n[0]);
        // (not with the release-only barriers of the staged mode: the rows a neighbour post-processes are
        //  read by plain loads)
        // (with at least grp + 1 rows per CTA the group range of a CTA lies inside the pass ranges of the CTA
        //  and its two neighbours: group boundaries sit within (-grp, padding + 1] of the pass boundaries)
        split = GS > 1 && !A.staged && L.n[0] / GS >= grp + 1;
        if (!split) {
            pr_lo[0] = r_lo[0];
            pr_hi[0] = r_hi[0];
            sl_lo[0] = r_lo[0] / L.R;
            sl_hi[0] = (r_hi[0] + L.R - 1) / L.R;
        }
    } else if (grp) {
        // (a CTA of activator rows: it takes part in the hand-over rounds of the others, without waiting)
        split = GS > 1 && !A.staged && L.n[0] / GS >= grp + 1;
    }
    // the CTA's pass rows are written: wait until the two neighbours have written theirs (same flavour as
    // group_sync: the fences make the neighbours' rows visible to plain loads)
    unsigned long long* nb_flags = A.bar + A.n_groups + (int64_t)gid * A.G;
    unsigned long long nb_epoch = 0;
    auto neighbour_sync = [&]() {
        __syncthreads();
        nb_epoch += 1;
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(nb_flags + g.cta), "l"(nb_epoch) : "memory");
            if (cS > 0) while (ld_acquire_u64(nb_flags + g.cta - 1) < nb_epoch) { }
            if (cS >= 0 && cS + 1 < GS) while (ld_acquire_u64(nb_flags + g.cta + 1) < nb_epoch) { }
            __threadfence();
        }
        __syncthreads();
    };
    // first row (type-local) and ion count of the CTA's group j
    auto chunk_rows = [&](int j, int& row0, int& m) {
        row0 = r_lo[0] + j * grp;
        m = min(grp, r_hi[0] - row0);
    };
    const int n_own_slices = (sl_hi[0] - sl_lo[0]) + (sl_hi[1] - sl_lo[1]);

    // ---- small lattices: make the CTA's own pair rows resident in shared memory ------------
    // Ownership is fixed for the whole launch and every system of the batch shares the lattice,
    // so the (partner, g) lines of the own slices are copied once; afterwards a Krylov iteration
    // touches global memory only for the (staged) gathered vector and the group barrier.
    if (A.resident) {
        double* sp = zs + N;                       // doubles first (alignment)
        double* g_s[2][2];
        for (int T = 0; T < 2; ++T)
            for (int U = 0; U < 2; ++U) {
                const PairBlock& B = L.blk[T][U];
                const int64_t base = B.n_slots ? B.line_ptr[sl_lo[T]] : 0;
                const int64_t cnt = B.n_slots ? B.line_ptr[sl_hi[T]] - base : 0;
                g_s[T][U] = sp;
                for (int64_t i = threadIdx.x; i < cnt * 32; i += STB_NT) sp[i] = B.g[0][base * 32 + i];
                sp += cnt * 32;
            }
        int64_t* lp_s[2][2];
        int64_t* ip = (int64_t*)sp;
        for (int T = 0; T < 2; ++T)
            for (int U = 0; U < 2; ++U) {
                const PairBlock& B = L.blk[T][U];
                const int cntp = sl_hi[T] - sl_lo[T] + 1;
                lp_s[T][U] = ip;
                if (B.n_slots)
                    for (int i = threadIdx.x; i < cntp; i += STB_NT) ip[i] = B.line_ptr[sl_lo[T] + i];
                ip += cntp;
            }
        uint32_t* cp = (uint32_t*)ip;
        uint32_t* c_s[2][2];
        for (int T = 0; T < 2; ++T)
            for (int U = 0; U < 2; ++U) {
                const PairBlock& B = L.blk[T][U];
                const int64_t base = B.n_slots ? B.line_ptr[sl_lo[T]] : 0;
                const int64_t cnt = B.n_slots ? B.line_ptr[sl_hi[T]] - base : 0;
                c_s[T][U] = cp;
                for (int64_t i = threadIdx.x; i < cnt * 32; i += STB_NT) cp[i] = B.col[base * 32 + i];
                cp += cnt * 32;
            }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int T = 0; T < 2; ++T)
                for (int U = 0; U < 2; ++U) {
                    PairBlock& B = L.blk[T][U];
                    if (!B.n_slots) continue;
                    const int64_t base = B.line_ptr[sl_lo[T]];
                    B.col = c_s[T][U] - base * 32;
                    B.g[0] = g_s[T][U] - base * 32;
                    B.line_ptr = lp_s[T][U] - sl_lo[T];
                }
        }
        __syncthreads();
    }

    double* W = A.work + (int64_t)gid * A.work_stride;
    double* D = W;                       // [8][N]
    double* yk = D + 8 * (int64_t)N;
    double* dd = yk + N;
    double* psi = dd + N;
    double* scale = psi + N;
    double* f = scale + N;
    double* wv = f + N;
    double* gv = wv + N;
    double* wbuf = gv + N;
    double* V = wbuf + N;                // [KRYLOV+1][N]
    double* F = V + (int64_t)(STB_KRYLOV + 1) * N;
    double* Minv = F + L.NF;
    double* Ginv = Minv + A.NMinv;       // [groups][col][row] inverses of the group blocks

    const double KAPPA[6] = {0.0, -0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0.0};
    double GAMMA[6], ALPHA[6], ERRC[6];
    GAMMA[0] = 0.0;
    for (int k = 1; k <= 5; ++k) GAMMA[k] = GAMMA[k - 1] + 1.0 / k;
    for (int k = 0; k <= 5; ++k) {
        ALPHA[k] = (1.0 - KAPPA[k]) * GAMMA[k];
        ERRC[k] = KAPPA[k] * GAMMA[k] + 1.0 / (k + 1);
    }
    const double sqrtN = sqrt((double)N);
    long long tick = clock64();
#define PROF(k)                                              \
    if (threadIdx.x == 0) {                                  \
        long long now_ = clock64();                          \
        C.prof[k] += now_ - tick;                            \
        tick = now_;                                         \
    }

    // ---- RHS phase: f = rhs(z) for own rows, fields F; optional block-Jacobi setup and
    //      preconditioned scaled residual (Newton) -------------------------------------------
    // mode 0: plain RHS (f only).  mode 1: Newton: also Minv(c), wv = Minv(c f - psi - d)/scale,
    //         gv = wv*scale; returns partial ||wv||^2 and a non-finite flag in part[].
    // copy the gathered vector (written by all CTAs of the group) through L2 into shared memory
    auto stage = [&](const double* z) -> const double* {
        if (!A.staged) return z;
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += STB_NT) zs[i] = __ldcg(z + i);
        __syncthreads();
        return zs;
    };
    const int n_own_A = sl_hi[1] - sl_lo[1];
    // ---- dense rows fed by the TMA engine (stb_tma.cuh): the CTA's own rows in tiles of <= 16,
    //      every warp reduces one row per tile and runs `epi` on it -----------------------------
    TmaRing R;
    if (A.tma) tma_ring_init(R, zs, A.tma_kc, A.tma_stages, 0);
    auto tma_rows = [&](const double* zg, double fscale, auto&& epi) {
        tma_rows_pass(R, L, pr_lo, pr_hi, 1, zg, fscale, warp, lane, epi);   // activator rows first
    };
    auto rhs_phase = [&](const double* z, double* fout, int mode, double c, double (&part)[2]) {
        part[0] = 0.0;
        part[1] = 0.0;
        const double* zg = stage(z);
        // per-ion epilogue of a row handled by W >= 16 lanes: lane j <-> state j of the ion
        auto rhs_row = [&](int T, int q, bool valid, int W, int gl, double fl) {
            const int s = L.s[T], nT = L.n[T];
            // cooperative (three-ion) process: one more field of the ion, summed over its triple list
            // by the whole warp (rows of such lattices are handled by W = 32 lanes, one row per warp)
            for (int r = 0; r < L.n_coop; ++r) {
                const CoopList& Cl = L.coop[r];
                if (Cl.T != T) continue;
                const double cf = coop_list_field_cg(Cl, valid ? q : nT - 1, z + L.off[Cl.T1] + Cl.col1 * L.n[Cl.T1],
                                                  z + L.off[Cl.T2] + Cl.col2 * L.n[Cl.T2], nullptr, nullptr, lane);
                if (gl == Cl.fslot) fl = cf;
            }
            const bool act = valid && gl < s;
            const int e = L.off[T] + (gl < s ? gl : 0) * nT + (valid ? q : nT - 1);
            const double yj = gl < s ? z[e] : 0.0;
            const double fj = coop_rhs(L, P, T, W, gl, yj, fl);
            if (act) fout[e] = fj;
            if (valid && gl < L.nf[T]) F[L.foff[T] + gl * nT + q] = fl;
            if (mode == 1 && grp && T == 0) {
                // group block: keep the ROW of M = I - cJ_qq and the raw residual; the block over the
                // whole group is assembled, inverted and applied once the CTA's rows are done
                double am[STB_MAXS];
                coop_block_row(L, P, T, W, gl, fl, c, am);
                double* Mp = Minv + q;
                if (act) {
#pragma unroll
                    for (int b = 0; b < STB_MAXS; ++b)
                        if (b < s) Mp[(int64_t)(gl * s + b) * nT] = am[b];
                    gv[e] = c * fj - psi[e] - dd[e];
                    if (!isfinite(fj)) part[1] = 1.0;
                }
            } else if (mode == 1) {
                double ai[STB_MAXS];
                coop_block_inverse(L, P, T, W, gl, fl, c, ai);
                double* Mp = Minv + (T ? L.s[0] * L.s[0] * L.n[0] : 0) + q;
                if (act) {
#pragma unroll
                    for (int b = 0; b < STB_MAXS; ++b)
                        if (b < s) Mp[(int64_t)(gl * s + b) * nT] = ai[b];
                }
                const double rj = act ? c * fj - psi[e] - dd[e] : 0.0;
                const double xy = coop_apply_rows(ai, s, W, rj);
                if (act) {
                    const double x = xy / scale[e];
                    wv[e] = x;
                    gv[e] = xy;
                    part[0] = fma(x, x, part[0]);
                    if (!isfinite(fj)) part[1] = 1.0;
                }
            }
        };
        // ---- group blocks (mode 1): B = I - cJ over the grp ions of a group = ion blocks (rows kept
        //      in Minv) + the S-S pair couplings inside the group at the iterate z (interaction donor
        //      q, partner k: rate = C g_qk y_q[ii] y_k[fi] moves q: ii -> if, k: fi -> ff), inverted in
        //      shared memory by the whole CTA (Gauss-Jordan without pivoting: column-diagonally-
        //      dominant M-matrix), stored and applied to the residual of the group
        //      Without the TMA ring: the whole CTA per group, one tile.
        auto group_blocks_cta = [&]() {
            const int s0 = L.s[0], n0 = L.n[0], GD = STB_GDIM, LD = STB_GDIM + 1;
            const double* g00 = L.blk[0][0].g[0];
            const double* yS = z + L.off[0];
            for (int gq = 0; gq < n_ch; ++gq) {
                int row0, mg;
                chunk_rows(gq, row0, mg);
                __syncthreads();
                for (int e = threadIdx.x; e < GD * GD; e += STB_NT) {
                    const int rho = e / GD, gam = e - rho * GD;
                    const int mq = rho / s0, aq = rho - mq * s0, mk = gam / s0, ak = gam - mk * s0;
                    const int iq = row0 + mq, ik = row0 + mk;
                    double v;
                    if (mq >= mg || mk >= mg) {
                        v = rho == gam ? 1.0 : 0.0;               // padding of a short group
                    } else if (mq == mk) {
                        v = Minv[(int64_t)(aq * s0 + ak) * n0 + iq];
                    } else {
                        v = 0.0;
                        const double gw = g00[(int64_t)iq * L.ld[0] + ik];
                        for (int p = 0; p < ss_n; ++p) {
                            const int p_ii = ss_st[p][0], p_fi = ss_st[p][1], p_if = ss_st[p][2], p_ff = ss_st[p][3];
                            const double cg = c * ss_C[p] * gw;
                            if (ak == p_fi) {            // q donor, k partner: d rate / d y_k[fi]
                                const double w = cg * yS[p_ii * n0 + iq];
                                if (aq == p_ii) v += w;
                                if (aq == p_if) v -= w;
                            }
                            if (ak == p_ii) {            // k donor, q partner: d rate / d y_k[ii]
                                const double w = cg * yS[p_fi * n0 + iq];
                                if (aq == p_fi) v += w;
                                if (aq == p_ff) v -= w;
                            }
                        }
                    }
                    Gs[rho * LD + gam] = v;
                }
                __syncthreads();
                for (int p = 0; p < GD; ++p) {
                    double nv[(GD * GD + STB_NT - 1) / STB_NT];
                    const double piv = Gs[p * LD + p];
                    const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
#pragma unroll
                    for (int u = 0; u < (GD * GD + STB_NT - 1) / STB_NT; ++u) {
                        const int e = threadIdx.x + u * STB_NT;
                        if (e < GD * GD) {
                            const int i = e / GD, j = e - i * GD;
                            const double aip = Gs[i * LD + p] * inv, apj = Gs[p * LD + j];
                            nv[u] = i == p ? (j == p ? inv : apj * inv)
                                           : (j == p ? -aip : fma(-aip, apj, Gs[i * LD + j]));
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < (GD * GD + STB_NT - 1) / STB_NT; ++u) {
                        const int e = threadIdx.x + u * STB_NT;
                        if (e < GD * GD) Gs[(e / GD) * LD + (e % GD)] = nv[u];
                    }
                    __syncthreads();
                }
                double* Gp = Ginv + (int64_t)(ch_slot + gq) * GD * GD;
                for (int e = threadIdx.x; e < GD * GD; e += STB_NT) {
                    const int col = e / GD, row = e - col * GD;
                    Gp[e] = Gs[row * LD + col];
                }
                if (warp == 0) {       // x = B^-1 r for the rows of the group, lane = row
                    const int m = lane / s0, a = lane - m * s0, iq = row0 + m;
                    const bool ok = m < mg;
                    const int e = L.off[0] + a * n0 + (ok ? iq : n0 - 1);
                    const double r = ok ? gv[e] : 0.0;
                    double acc = 0.0;
                    for (int col = 0; col < GD; ++col)
                        acc = fma(Gs[lane * LD + col], __shfl_sync(FULL, r, col), acc);
                    if (ok) {
                        const double x = acc / scale[e];
                        wv[e] = x;
                        gv[e] = acc;
                        part[0] = fma(x, x, part[0]);
                    }
                }
            }
            __syncthreads();
        };
        //      TMA passes: one warp per group, lane = row, tile in shared memory - the ring is idle between
        //      two passes and holds a tile per warp
        auto group_blocks_warp = [&]() {
            const int s0 = L.s[0], n0 = L.n[0], GD = STB_GDIM, LD = STB_GDIM + 1;
            const double* g00 = L.blk[0][0].g[0];
            const double* yS = z + L.off[0];
            const int nt = min(STB_NW, (int)(tma_ring_bytes(A.tma_kc, A.tma_stages) / (GD * LD * sizeof(double))));
            __syncthreads();
            if (warp < nt) {
                double* Tq = R.ring + warp * GD * LD;
                const int mq = lane / s0, aq = lane - mq * s0;
                for (int gq = warp; gq < n_ch; gq += nt) {
                    int row0, mg;
                    chunk_rows(gq, row0, mg);
                    const bool ok = mq < mg;
                    const int iq = ok ? row0 + mq : row0;
                    // row `lane` of the block: identity / ion block first, then the couplings to the other
                    // ions of the group (their pair rates are loaded eight at a time: independent loads)
                    for (int gam = 0; gam < GD; ++gam) Tq[lane * LD + gam] = lane == gam ? 1.0 : 0.0;
                    double yq[8];
#pragma unroll
                    for (int a = 0; a < 8; ++a) yq[a] = (ok && a < s0) ? yS[a * n0 + iq] : 0.0;
                    if (ok) {
                        for (int ak = 0; ak < s0; ++ak)
                            Tq[lane * LD + mq * s0 + ak] = Minv[(int64_t)(aq * s0 + ak) * n0 + iq];
                        for (int mk0 = 0; mk0 < mg; mk0 += 8) {
                            double gw[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int mk = mk0 + u;
                                gw[u] = (mk < mg && mk != mq) ? g00[(int64_t)iq * L.ld[0] + row0 + mk] : 0.0;
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int mk = mk0 + u;
                                if (mk >= mg || mk == mq) continue;
                                for (int ak = 0; ak < s0; ++ak) {
                                    double v = 0.0;
                                    for (int p = 0; p < ss_n; ++p) {
                                        const int p_ii = ss_st[p][0], p_fi = ss_st[p][1], p_if = ss_st[p][2], p_ff = ss_st[p][3];
                                        const double cg = c * ss_C[p] * gw[u];
                                        if (ak == p_fi) {            // q donor, k partner: d rate / d y_k[fi]
                                            const double w = cg * yq[p_ii];
                                            if (aq == p_ii) v += w;
                                            if (aq == p_if) v -= w;
                                        }
                                        if (ak == p_ii) {            // k donor, q partner: d rate / d y_k[ii]
                                            const double w = cg * yq[p_fi];
                                            if (aq == p_fi) v += w;
                                            if (aq == p_ff) v -= w;
                                        }
                                    }
                                    Tq[lane * LD + mk * s0 + ak] = v;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    // in-place Gauss-Jordan: the pivot row is scaled by the lanes (one column each), then
                    // every other lane eliminates the pivot column from its own row
                    for (int p = 0; p < GD; ++p) {
                        const double piv = Tq[p * LD + p];
                        const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
                        const double fr = Tq[lane * LD + p];
                        __syncwarp();
                        Tq[p * LD + lane] = lane == p ? inv : Tq[p * LD + lane] * inv;
                        __syncwarp();
                        if (lane != p) {
                            // eight columns at a time: the loads of a chunk are independent of its stores
                            for (int j0 = 0; j0 < GD; j0 += 8) {
                                double pj[8], ow[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u) {
                                    pj[u] = Tq[p * LD + j0 + u];
                                    ow[u] = Tq[lane * LD + j0 + u];
                                }
#pragma unroll
                                for (int u = 0; u < 8; ++u)
                                    Tq[lane * LD + j0 + u] = j0 + u == p ? -fr * pj[u] : fma(-fr, pj[u], ow[u]);
                            }
                        }
                        __syncwarp();
                    }
                    double* Gp = Ginv + (int64_t)(ch_slot + gq) * GD * GD;
                    const int e = L.off[0] + aq * n0 + (ok ? iq : n0 - 1);
                    const double r = ok ? gv[e] : 0.0;
                    double acc = 0.0;
                    for (int col = 0; col < GD; ++col) {
                        const double bv = Tq[lane * LD + col];
                        Gp[col * GD + lane] = bv;
                        acc = fma(bv, __shfl_sync(FULL, r, col), acc);      // x = B^-1 r, lane = row
                    }
                    if (ok) {
                        const double x = acc / scale[e];
                        wv[e] = x;
                        gv[e] = acc;
                        part[0] = fma(x, x, part[0]);
                    }
                    __syncwarp();
                }
            }
            // the next pass writes the ring through the async proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
        };
        // lagged blocks (CVODE's rule, as the cluster integrator): between refreshes the stored inverses
        // of an earlier step precondition the residual - B~^-1 (M v - c k(v)) stays exact in M and k
        auto group_blocks_stored = [&]() {
            __syncthreads();
            const int s0 = L.s[0], n0 = L.n[0], GD = STB_GDIM;
            for (int gq = warp; gq < n_ch; gq += STB_NW) {
                int row0, mg;
                chunk_rows(gq, row0, mg);
                const int m = lane / s0, a = lane - m * s0, iq = row0 + m;
                const bool ok = m < mg;
                const int e = L.off[0] + a * n0 + (ok ? iq : n0 - 1);
                const double r = ok ? gv[e] : 0.0;
                const double* Gp = Ginv + (int64_t)(ch_slot + gq) * GD * GD + lane;
                double acc = 0.0;
#pragma unroll 8
                for (int col = 0; col < GD; ++col)
                    acc = fma(Gp[col * GD], __shfl_sync(FULL, r, col), acc);
                if (ok) {
                    const double x = acc / scale[e];
                    wv[e] = x;
                    gv[e] = acc;
                    part[0] = fma(x, x, part[0]);
                }
            }
            __syncthreads();
        };
        auto group_blocks = [&]() {
            if (!C.pc_refresh) group_blocks_stored();
            else if (A.tma) group_blocks_warp();
            else group_blocks_cta();
        };
        if (A.tma) {
            tma_rows(zg, 1.0, rhs_row);
            if (split) neighbour_sync();
            if (mode == 1 && grp) group_blocks();
            return;
        }
        for (int task = warp; task < n_own_slices; task += STB_NW) {
            int T = task < n_own_A;          // activator rows first (the expensive epilogues)
            int slice = T ? sl_lo[1] + task : sl_lo[0] + task - n_own_A;
            if (L.lpr >= 16) {      // warp-cooperative epilogue: lane j <-> state j of the ion
                const int W = L.lpr, gl = lane & (W - 1);
                const double fl = A.resident ? slice_fields_resident(L, T, slice, zg, lane, 1.0, nullptr)
                                             : slice_fields(L, T, slice, zg, lane, 1.0, nullptr);
                const int q = slice * L.R + lane / W;
                rhs_row(T, q, q < L.n[T], W, gl, fl);
                continue;
            }
            double Fq[2 * STB_MAXC];
            if (A.resident) slice_fields_resident(L, T, slice, zg, lane, 1.0, Fq);
            else slice_fields(L, T, slice, zg, lane, 1.0, Fq);
            int q = slice * L.R + lane / L.lpr;
            if ((lane % L.lpr) == 0 && q < L.n[T]) {
                double zq[STB_MAXS], oq[STB_MAXS];
                load_row(L, T, q, z, zq);
                row_rhs(L, P, T, zq, Fq, oq);
                store_row(L, T, q, fout, oq);
                const int nT = L.n[T], s = L.s[T];
                double* Fp = F + L.foff[T] + q;
                for (int k = 0; k < L.nf[T]; ++k) Fp[(int64_t)k * nT] = Fq[k];
                if (mode == 1) {
                    double Am[STB_MAXS * STB_MAXS], Ai[STB_MAXS * STB_MAXS];
                    row_block_inverse(L, P, T, Fq, c, Am, Ai);
                    double* Mp = Minv + (T ? L.s[0] * L.s[0] * L.n[0] : 0) + q;
                    for (int k = 0; k < s * s; ++k) Mp[(int64_t)k * nT] = Ai[k];
                    const int base = L.off[T] + q;
                    double r[STB_MAXS];
                    for (int a = 0; a < s; ++a) {
                        int e = base + a * nT;
                        r[a] = c * oq[a] - psi[e] - dd[e];
                        if (!isfinite(oq[a])) part[1] = 1.0;
                    }
                    for (int a = 0; a < s; ++a) {
                        double acc = 0.0;
                        for (int b = 0; b < s; ++b) acc = fma(Ai[a * s + b], r[b], acc);
                        int e = base + a * nT;
                        double x = acc / scale[e];
                        wv[e] = x;
                        gv[e] = acc;
                        part[0] = fma(x, x, part[0]);
                    }
                }
            }
        }
        if (split) neighbour_sync();
        if (mode == 1 && grp) group_blocks();
    };

    // ---- Arnoldi phase: wbuf = Minv (v - c J v) / scale with v = inrm * gv ------------------
    auto arnoldi_phase = [&](double inrm, double c) {
        const double* zg = stage(gv);
        // M^-1 (I - cJ) v = v - c Minv k,  k = coupling of the ion to its partners' v
        auto arn_row = [&](int T, int q, bool valid, int W, int gl, double fvl) {
            const int s = L.s[T], nT = L.n[T];
            const bool act = valid && gl < s;
            const int qc = valid ? q : nT - 1;
            // cooperative field linearised at the iterate yk in the direction inrm * gv
            for (int r = 0; r < L.n_coop; ++r) {
                const CoopList& Cl = L.coop[r];
                if (Cl.T != T) continue;
                const int o1 = L.off[Cl.T1] + Cl.col1 * L.n[Cl.T1], o2 = L.off[Cl.T2] + Cl.col2 * L.n[Cl.T2];
                const double cf = inrm * coop_list_field_cg(Cl, qc, yk + o1, yk + o2, gv + o1, gv + o2, lane);
                if (gl == Cl.fslot) fvl = cf;
            }
            const int e = L.off[T] + (gl < s ? gl : 0) * nT + qc;
            const double yj = gl < s ? yk[e] : 0.0;
            const double kj = coop_coupling(L, P, T, W, gl, yj, fvl);
            const double* Mp = Minv + (T ? L.s[0] * L.s[0] * L.n[0] : 0) + qc;
            const int row = gl < s ? gl : 0;
            if (grp && T == 0) {
                // group block: B^-1 (I - cJ) v = B^-1 (M v - c k(v)) with the rows of the ion's own
                // M = I - cJ_qq (kept in Minv); B^-1 is applied when the CTA's rows are done
                const double vj = gl < s ? inrm * gv[e] : 0.0;
                double mv = 0.0;
                for (int b = 0; b < s; ++b)
                    mv = fma(Mp[(int64_t)(row * s + b) * nT], __shfl_sync(FULL, vj, b, W), mv);
                if (act) wbuf[e] = mv - c * kj;
                return;
            }
            double mk = 0.0;
            for (int b = 0; b < s; ++b)
                mk = fma(Mp[(int64_t)(row * s + b) * nT], __shfl_sync(FULL, kj, b, W), mk);
            if (act) wbuf[e] = (inrm * gv[e] - c * mk) / scale[e];
        };
        // wbuf <- B^-1 wbuf / scale over the CTA's own groups, one warp per group, lane = row
        auto group_apply = [&]() {
            __syncthreads();
            const int s0 = L.s[0], n0 = L.n[0], GD = STB_GDIM;
            for (int gq = warp; gq < n_ch; gq += STB_NW) {
                int row0, mg;
                chunk_rows(gq, row0, mg);
                const int m = lane / s0, a = lane - m * s0, iq = row0 + m;
                const bool ok = m < mg;
                const int e = L.off[0] + a * n0 + (ok ? iq : n0 - 1);
                const double r = ok ? wbuf[e] : 0.0;
                const double* Gp = Ginv + (int64_t)(ch_slot + gq) * GD * GD + lane;
                double acc = 0.0;
#pragma unroll 8
                for (int col = 0; col < GD; ++col)
                    acc = fma(Gp[col * GD], __shfl_sync(FULL, r, col), acc);     // (written by this CTA only)
                if (ok) wbuf[e] = acc / scale[e];
            }
        };
        if (A.tma) {
            long long t0_ = clock64();
            tma_rows(zg, inrm, arn_row);
            if (threadIdx.x == 0) C.prof[6] += clock64() - t0_;
            if (split) neighbour_sync();
            if (grp) group_apply();
            return;
        }
        for (int task = warp; task < n_own_slices; task += STB_NW) {
            int T = task < n_own_A;
            int slice = T ? sl_lo[1] + task : sl_lo[0] + task - n_own_A;
            if (L.lpr >= 16) {
                const int W = L.lpr, gl = lane & (W - 1);
                long long t0_ = clock64();
                const double fvl = A.resident ? slice_fields_resident(L, T, slice, zg, lane, inrm, nullptr)
                                               : slice_fields(L, T, slice, zg, lane, inrm, nullptr);
                if (threadIdx.x == 0) C.prof[6] += clock64() - t0_;
                const int q = slice * L.R + lane / W;
                arn_row(T, q, q < L.n[T], W, gl, fvl);
                continue;
            }
            double Fv[2 * STB_MAXC];
            if (A.resident) slice_fields_resident(L, T, slice, zg, lane, inrm, Fv);
            else slice_fields(L, T, slice, zg, lane, inrm, Fv);
            int q = slice * L.R + lane / L.lpr;
            if ((lane % L.lpr) == 0 && q < L.n[T]) {
                const int nT = L.n[T], s = L.s[T];
                double Fq[2 * STB_MAXC], yq[STB_MAXS], vq[STB_MAXS], jq[STB_MAXS];
                const double* Fp = F + L.foff[T] + q;
                for (int k = 0; k < L.nf[T]; ++k) Fq[k] = Fp[(int64_t)k * nT];
                load_row(L, T, q, yk, yq);
                load_row(L, T, q, gv, vq);
                for (int a = 0; a < s; ++a) vq[a] *= inrm;
                row_jv(L, P, T, yq, vq, Fq, Fv, jq);
                for (int a = 0; a < s; ++a) jq[a] = vq[a] - c * jq[a];
                const double* Mp = Minv + (T ? L.s[0] * L.s[0] * L.n[0] : 0) + q;
                const int base = L.off[T] + q;
                for (int a = 0; a < s; ++a) {
                    double acc = 0.0;
                    for (int b = 0; b < s; ++b) acc = fma(Mp[(int64_t)(a * s + b) * nT], jq[b], acc);
                    int e = base + a * nT;
                    wbuf[e] = acc / scale[e];
                }
            }
        }
        if (split) neighbour_sync();
        if (grp) group_apply();
    };

    // dots of `vec` (own elements) against V[0..nv) and itself: rv[i] (i<nv), rv[nv] = <vec,vec>
    auto own_dots = [&](const double* vec, int nv) {
        __syncthreads();
        for (int i = warp; i <= nv; i += STB_NW) {
            const double* Vi = i < nv ? V + (int64_t)i * N : vec;
            double acc = 0.0;
            for (int T = 0; T < 2; ++T) {
                const int cnt = r_hi[T] - r_lo[T];
                const int tot = cnt * L.s[T];
                for (int k = lane; k < tot; k += 32) {
                    int e = L.off[T] + (k / cnt) * L.n[T] + r_lo[T] + (k % cnt);
                    acc = fma(vec[e], Vi[e], acc);
                }
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) rv[i] = acc;
        }
        __syncthreads();
    };

    // D[0..order] <- RU^T D[0..order] for own elements (C.RU set by thread 0)
    auto apply_change_D = [&](int order) {
        OWN_LOOP(e)
            double dv[6], nd[6];
            for (int j = 0; j <= order; ++j) dv[j] = D[(int64_t)j * N + e];
            for (int i = 0; i <= order; ++i) {
                double acc = 0.0;
                for (int j = 0; j <= order; ++j) acc = fma(C.RU[j * 6 + i], dv[j], acc);
                nd[i] = acc;
            }
            for (int i = 0; i <= order; ++i) D[(int64_t)i * N + e] = nd[i];
        OWN_END
    };

    for (int sysl = gid; sysl < A.n_sys; sysl += A.n_groups) {
        const int sys = A.first + sysl;
        __syncthreads();
        load_sys_params(L, A.blocks, A.strengths, sys, A.pulse, P);
        const double* y0 = A.y0 + (int64_t)sysl * N;
        double* y_out = A.y_out + (int64_t)sysl * A.n_t * N;
        const double rtol = A.rtol, atol = A.atol;
        if (grp) {
            __syncthreads();
            if (threadIdx.x == 0) {
                int m = 0;
                for (int p = 0; p < L.n_proc; ++p) {
                    const stb_process pr = A.procs[p];
                    if (pr.donor_type != 0 || pr.partner_type != 0 || P.C[p] == 0.0) continue;
                    ss_st[m][0] = pr.ii; ss_st[m][1] = pr.fi; ss_st[m][2] = pr.if_; ss_st[m][3] = pr.ff;
                    ss_C[m] = P.C[p];
                    ++m;
                }
                ss_n = m;
            }
        }
        if (threadIdx.x == 0) {
            C.t = A.t[0];
            C.t_end = A.t[A.n_t - 1];
            C.order = 1;
            C.n_equal = 0;
            C.status = 0;
            C.k_out = 1;
            C.steps_since_out = 0;
            C.newton_tol = fmax(10.0 * 2.220446049250313e-16 / rtol, fmin(0.03, sqrt(rtol)));
            C.tol_lin = A.lin_rms * sqrtN;
            C.lin_tight = 0;
            C.pc_age = -1;
            C.pc_refresh = 1;
            C.crate = 1.0;
            C.since_verify = 0;
            C.gfac = 0.0;
            memset(&C.st, 0, sizeof(stb_stats));
            for (int k = 0; k < 8; ++k) C.prof[k] = 0;
            tick = clock64();
        }
        // D0 = y0, yk = y0, row 0 of the output = y0 verbatim (odesolver.py:107)
        OWN_LOOP(e)
            double v = y0[e];
            D[e] = v;
            yk[e] = v;
            y_out[L.ref_of_plane[e]] = v;
            for (int j = 1; j < 8; ++j) D[(int64_t)j * N + e] = 0.0;
        OWN_END
        group_sync(g);
        double part[2];
        rhs_phase(yk, f, 0, 0.0, part);
        if (threadIdx.x == 0) C.st.n_rhs += 1;
        // ---- initial step (Hairer-Wanner, as scipy select_initial_step) -------------------
        {
            double v3[3] = {0.0, 0.0, 0.0};
            __syncthreads();
            OWN_LOOP(e)
                double sc = atol + rtol * fabs(yk[e]);
                scale[e] = sc;
                double a = yk[e] / sc, b = f[e] / sc;
                v3[0] = fma(a, a, v3[0]);
                v3[1] = fma(b, b, v3[1]);
                if (!isfinite(f[e])) v3[2] = 1.0;
            OWN_END
            block_sum<3>(v3, red, rv);
            group_reduce(g, rv, 3);
            double d0 = sqrt(rv[0] / N), d1 = sqrt(rv[1] / N);
            double bad = rv[2];
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            __syncthreads();
            OWN_LOOP(e)
                gv[e] = yk[e] + h0 * f[e];
            OWN_END
            group_sync(g);
            rhs_phase(gv, wbuf, 0, 0.0, part);
            double v1[1] = {0.0};
            __syncthreads();
            OWN_LOOP(e)
                double a = (wbuf[e] - f[e]) / scale[e];
                v1[0] = fma(a, a, v1[0]);
            OWN_END
            block_sum<1>(v1, red, rv);
            group_reduce(g, rv, 1);
            if (threadIdx.x == 0) {
                C.st.n_rhs += 1;
                double d2 = sqrt(rv[0] / N) / h0;
                double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3)
                                                         : sqrt(0.01 / fmax(d1, d2));
                C.h = fmin(fmin(100.0 * h0, h1), C.t_end - C.t);
                if (bad != 0.0 || !isfinite(C.h) || !(C.h > 0.0)) C.status = 3;
            }
            __syncthreads();
            // the fields F must correspond to yk again before the first Newton iteration: they
            // are recomputed by the first RHS of the step.  D1 = f0 * h
            const double h = C.h;
            OWN_LOOP(e)
                D[(int64_t)N + e] = f[e] * h;
            OWN_END
        }
        __syncthreads();

        // =============================== time stepping =====================================
        while (true) {
            __syncthreads();
            if (C.status != 0 || C.k_out >= A.n_t) break;
            // ---- one step attempt ------------------------------------------------------------
            if (threadIdx.x == 0) {
                C.flag = 0;
                double t_new = C.t + C.h;
                if (t_new > C.t_end && (t_new - C.t_end) > 1e-14 * fabs(C.t_end)) {
                    double factor = (C.t_end - C.t) / C.h;
                    compute_RU(C.order, factor, C.RU);
                    C.h = C.t_end - C.t;
                    C.n_equal = 0;
                    C.flag = 1;
                    C.gfac *= stb_ipow(factor, C.order + 1);
                } else if (t_new > C.t_end) {
                    C.h = C.t_end - C.t;  // round-off sized overshoot: land exactly on t_end
                }
                double eps_t = 10.0 * fabs(nextafter(C.t, 1e300) - C.t);
                if (C.h < eps_t) C.status = 2;
                if (C.steps_since_out >= A.nsteps) C.status = 1;
                C.c = C.h / ALPHA[C.order];
            }
            __syncthreads();
            if (C.status != 0) break;
            if (C.flag) apply_change_D(C.order);
            __syncthreads();
            const int order = C.order;
            const double c = C.c;
            const double gf = A.guess ? C.gfac : 0.0;
            // predictor, scale, psi; Newton starts from the predictor plus the previous step's
            // correction (D[order + 1]) rescaled to this step size
            OWN_LOOP(e)
                double yp = 0.0, ps = 0.0;
                for (int j = 0; j <= order; ++j) {
                    double dj = D[(int64_t)j * N + e];
                    yp += dj;
                    if (j >= 1) ps = fma(GAMMA[j], dj, ps);
                }
                const double d0 = gf != 0.0 ? gf * D[(int64_t)(order + 1) * N + e] : 0.0;
                yk[e] = yp + d0;
                scale[e] = atol + rtol * fabs(yp);
                psi[e] = ps / ALPHA[order];
                dd[e] = d0;
            OWN_END
            // ---- Newton-Krylov ----------------------------------------------------------------
            if (threadIdx.x == 0) { C.newton_it = 0; C.flag = 0; C.dy_norm_old = -1.0; }
            __syncthreads();
            while (true) {       // flag: 0 iterate, 1 converged, 2 failed
                PROF(7)
                if (threadIdx.x == 0) {
                    // refresh the group blocks when c = h / alpha has drifted by 30 %, after STB_PC_MAXAGE
                    // steps, or after a failed Newton iteration (identical decision in every CTA)
                    C.pc_refresh = !A.lag || C.pc_age < 0 || C.pc_age >= STB_PC_MAXAGE ||
                                   fabs(c / C.pc_c - 1.0) > STB_PC_DRIFT;
                    if (C.pc_refresh) { C.pc_c = c; C.pc_age = 0; }
                }
                group_sync(g);   // yk complete in the whole group
                PROF(3)
                rhs_phase(yk, f, 1, c, part);
                PROF(0)
                double v2[2] = {part[0], part[1]};
                block_sum<2>(v2, red, rv);
                group_reduce(g, rv, 2);
                PROF(3)
                if (threadIdx.x == 0) {
                    C.st.n_rhs += 1;
                    C.st.n_lin += 1;
                    C.beta = sqrt(rv[0]);
                    C.kk = 0;
                    if (rv[1] != 0.0 || !isfinite(C.beta)) C.flag = 2;
                    gg[0] = C.beta;
                    C.inrm = C.beta > 0.0 ? 1.0 / C.beta : 0.0;
                }
                __syncthreads();
                if (C.flag == 2) break;
                // ---- GMRES ----------------------------------------------------------------
                bool lin_done = !(C.beta > C.tol_lin);
                int j = 0;
                while (!lin_done) {
                    const double inrm = C.inrm;
                    // V_j = inrm * wv (own elements); then the fused operator pass
                    double* Vj = V + (int64_t)j * N;
                    PROF(7)
                    OWN_LOOP(e)
                        Vj[e] = inrm * wv[e];
                    OWN_END
                    PROF(4)
                    arnoldi_phase(inrm, c);
                    PROF(1)
                    own_dots(wbuf, j + 1);
                    PROF(2)
                    group_reduce(g, rv, j + 2);
                    PROF(3)
                    const double ww = rv[j + 1];
                    if (threadIdx.x <= (unsigned)j) H[threadIdx.x * STB_KRYLOV + j] = rv[threadIdx.x];
                    __syncthreads();
                    double v1[1] = {0.0};
                    OWN_LOOP(e)
                        double acc = wbuf[e];
                        for (int i = 0; i <= j; ++i)
                            acc = fma(-H[i * STB_KRYLOV + j], V[(int64_t)i * N + e], acc);
                        wv[e] = acc;
                        gv[e] = acc * scale[e];
                        v1[0] = fma(acc, acc, v1[0]);
                    OWN_END
                    PROF(4)
                    block_sum<1>(v1, red, rv);
                    group_reduce(g, rv, 1);
                    PROF(3)
                    double hn2 = rv[0];
                    if (hn2 < 0.01 * ww && hn2 > 0.0) {
                        // severe cancellation: one re-orthogonalisation pass (DGKS)
                        own_dots(wv, j + 1);
                        group_reduce(g, rv, j + 2);
                        if (threadIdx.x <= (unsigned)j) H[threadIdx.x * STB_KRYLOV + j] += rv[threadIdx.x];
                        double corr[STB_KRYLOV];
                        for (int i = 0; i <= j; ++i) corr[i] = rv[i];
                        __syncthreads();
                        v1[0] = 0.0;
                        OWN_LOOP(e)
                            double acc = wv[e];
                            for (int i = 0; i <= j; ++i)
                                acc = fma(-corr[i], V[(int64_t)i * N + e], acc);
                            wv[e] = acc;
                            gv[e] = acc * scale[e];
                            v1[0] = fma(acc, acc, v1[0]);
                        OWN_END
                        block_sum<1>(v1, red, rv);
                        group_reduce(g, rv, 1);
                        hn2 = rv[0];
                    }
                    if (threadIdx.x == 0) {
                        C.st.n_jv += 1;
                        double hn = sqrt(hn2);
                        double* Hc = H + j;   // column j, stride STB_KRYLOV
                        for (int i = 0; i < j; ++i) {
                            double a = Hc[i * STB_KRYLOV], b = Hc[(i + 1) * STB_KRYLOV];
                            Hc[i * STB_KRYLOV] = cs[i] * a + sn[i] * b;
                            Hc[(i + 1) * STB_KRYLOV] = -sn[i] * a + cs[i] * b;
                        }
                        double a = Hc[j * STB_KRYLOV];
                        double dnm = hypot(a, hn);
                        if (dnm == 0.0) { cs[j] = 1.0; sn[j] = 0.0; }
                        else { cs[j] = a / dnm; sn[j] = hn / dnm; }
                        Hc[j * STB_KRYLOV] = dnm;
                        gg[j + 1] = -sn[j] * gg[j];
                        gg[j] = cs[j] * gg[j];
                        C.kk = j + 1;
                        C.inrm = hn > 0.0 ? 1.0 / hn : 0.0;
                        C.beta = fabs(gg[j + 1]);
                        if (!(hn > 0.0) || !isfinite(hn)) C.beta = 0.0;
                    }
                    __syncthreads();
                    PROF(5)
                    ++j;
                    lin_done = !(C.beta > C.tol_lin) || j >= STB_KRYLOV;
                }
                // ---- x = V y, dy = x * scale; Newton control --------------------------------
                const int kk = C.kk;
                if (threadIdx.x == 0) {
                    for (int i = kk - 1; i >= 0; --i) {
                        double acc = gg[i];
                        for (int k = i + 1; k < kk; ++k) acc -= H[i * STB_KRYLOV + k] * yk_s[k];
                        double hd = H[i * STB_KRYLOV + i];
                        yk_s[i] = hd != 0.0 ? acc / hd : 0.0;
                    }
                }
                __syncthreads();
                double v1[1] = {0.0};
                OWN_LOOP(e)
                    double x = 0.0;
                    for (int i = 0; i < kk; ++i) x = fma(yk_s[i], V[(int64_t)i * N + e], x);
                    wbuf[e] = x;      // scaled Newton update
                    v1[0] = fma(x, x, v1[0]);
                OWN_END
                block_sum<1>(v1, red, rv);
                group_reduce(g, rv, 1);
                if (threadIdx.x == 0) {
                    C.flag = newton_decide(C, sqrt(rv[0] / N));     // 2: do not apply the update
                }
                __syncthreads();
                const int flag = C.flag;
                if (flag != 2) {
                    OWN_LOOP(e)
                        double dy = wbuf[e] * scale[e];
                        yk[e] += dy;
                        dd[e] += dy;
                    OWN_END
                }
                if (flag != 0) break;
                if (C.newton_it >= STB_NEWTON_MAXIT) {
                    __syncthreads();
                    if (threadIdx.x == 0) C.flag = 2;
                    __syncthreads();
                    break;
                }
            }
            __syncthreads();
            if (C.flag == 2) {
                // Newton / Krylov failure: halve the step (scipy: factor 0.5 with current Jacobian)
                if (threadIdx.x == 0) {
                    C.st.n_newton_fail += 1;
                    C.pc_age = -1;
                    C.lin_tight = A.lin_tight_steps;
                    C.tol_lin = fmin(A.lin_rms, STB_LIN_RMS) * sqrtN;
                    compute_RU(C.order, 0.5, C.RU);
                    C.h *= 0.5;
                    C.n_equal = 0;
                    C.steps_since_out += 1;
                    C.gfac = 0.0;
                }
                __syncthreads();
                apply_change_D(order);
                continue;
            }
            // ---- error test --------------------------------------------------------------------
            {
                double v1[1] = {0.0};
                __syncthreads();
                OWN_LOOP(e)
                    double sc = atol + rtol * fabs(yk[e]);
                    scale[e] = sc;
                    double er = ERRC[order] * dd[e] / sc;
                    v1[0] = fma(er, er, v1[0]);
                OWN_END
                block_sum<1>(v1, red, rv);
                group_reduce(g, rv, 1);
                if (threadIdx.x == 0) {
                    C.err_norm = sqrt(rv[0] / N);
                    C.safety = 0.9 * (2 * STB_NEWTON_MAXIT + 1) / (2.0 * STB_NEWTON_MAXIT + C.n_iter);
                    C.flag = 0;
                    if (C.err_norm > 1.0 || !isfinite(C.err_norm)) {
                        double factor = isfinite(C.err_norm)
                            ? fmax(STB_MIN_FACTOR, C.safety * pow(C.err_norm, -1.0 / (order + 1)))
                            : STB_MIN_FACTOR;
                        compute_RU(order, factor, C.RU);
                        C.h *= factor;
                        C.n_equal = 0;
                        C.st.n_rejected += 1;
                        C.lin_tight = A.lin_tight_steps;
                        C.tol_lin = fmin(A.lin_rms, STB_LIN_RMS) * sqrtN;
                        C.steps_since_out += 1;
                        C.flag = 1;
                        C.gfac *= stb_ipow(factor, order + 1);
                    }
                }
                __syncthreads();
                if (C.flag == 1) {
                    apply_change_D(order);
                    continue;
                }
            }
            // ---- accepted: update differences, dense output, order / step selection -------------
            const double h = C.h, t_new = C.t + C.h;
            {
                double v2[2] = {0.0, 0.0};
                __syncthreads();
                OWN_LOOP(e)
                    double dv = dd[e];
                    double dp1 = D[(int64_t)(order + 1) * N + e];
                    double d2 = dv - dp1;
                    D[(int64_t)(order + 2) * N + e] = d2;
                    D[(int64_t)(order + 1) * N + e] = dv;
                    double run = dv;
                    double dord = 0.0;
                    for (int i = order; i >= 0; --i) {
                        run += D[(int64_t)i * N + e];
                        D[(int64_t)i * N + e] = run;
                        if (i == order) dord = run;
                    }
                    double sc = scale[e];
                    if (order > 1) { double em = ERRC[order - 1] * dord / sc; v2[0] = fma(em, em, v2[0]); }
                    if (order < STB_MAXORD) { double ep = ERRC[order + 1] * d2 / sc; v2[1] = fma(ep, ep, v2[1]); }
                OWN_END
                // dense output for every requested time in (t, t_new]
                __syncthreads();
                while (true) {
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        C.flag = 0;
                        if (C.k_out < A.n_t) {
                            double tq = A.t[C.k_out];
                            if (tq <= t_new * (1.0 + 1e-15) || tq <= t_new) {
                                C.flag = tq >= t_new ? 2 : 1;
                                double pr = 1.0;
                                for (int jj = 0; jj < order; ++jj) {
                                    pr *= (tq - (t_new - h * jj)) / (h * (1 + jj));
                                    C.p[jj + 1] = pr;
                                }
                            }
                        }
                    }
                    __syncthreads();
                    if (C.flag == 0) break;
                    double* row = y_out + C.k_out * (int64_t)N;
                    const bool exact = C.flag == 2;
                    OWN_LOOP(e)
                        double acc = D[e];
                        if (!exact)
                            for (int jj = 1; jj <= order; ++jj)
                                acc = fma(C.p[jj], D[(int64_t)jj * N + e], acc);
                        row[L.ref_of_plane[e]] = acc;
                    OWN_END
                    __syncthreads();
                    if (threadIdx.x == 0) { C.k_out += 1; C.st.n_out = C.k_out; C.steps_since_out = 0; }
                }
                const bool need_order = (C.n_equal + 1) >= order + 1;
                if (need_order) {
                    block_sum<2>(v2, red, rv);
                    group_reduce(g, rv, 2);
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    C.st.n_steps += 1;
                    if (C.pc_age >= 0) C.pc_age += 1;
                    if (C.lin_tight > 0 && --C.lin_tight == 0) C.tol_lin = A.lin_rms * sqrtN;
                    C.steps_since_out += 1;
                    C.n_equal += 1;
                    C.t = t_new;
                    C.flag = 0;
                    C.gfac = 1.0;
                    if (need_order) {
                        double err_m = order > 1 ? sqrt(rv[0] / N) : INFINITY;
                        double err_p = order < STB_MAXORD ? sqrt(rv[1] / N) : INFINITY;
                        double fm = err_m > 0.0 ? pow(err_m, -1.0 / order) : INFINITY;
                        double f0 = C.err_norm > 0.0 ? pow(C.err_norm, -1.0 / (order + 1)) : INFINITY;
                        double fp = err_p > 0.0 ? pow(err_p, -1.0 / (order + 2)) : INFINITY;
                        if (order == 1) fm = 0.0;
                        if (order == STB_MAXORD) fp = 0.0;
                        if (isinf(err_m)) fm = 0.0;
                        if (isinf(err_p)) fp = 0.0;
                        int delta = 0;
                        double best = f0;
                        if (fm > best) { best = fm; delta = -1; }
                        if (fp > best) { best = fp; delta = 1; }
                        // numpy argmax picks the first maximum: order-1 wins ties, then same order
                        if (fm >= f0 && fm >= fp) { best = fm; delta = -1; }
                        else if (f0 >= fp) { best = f0; delta = 0; }
                        else { best = fp; delta = 1; }
                        C.order = order + delta;
                        double factor = fmin(STB_MAX_FACTOR, C.safety * best);
                        compute_RU(C.order, factor, C.RU);
                        C.h *= factor;
                        C.n_equal = 0;
                        C.flag = 1;
                        C.gfac = delta == 0 ? stb_ipow(factor, order + 1) : 0.0;
                    }
                    C.st.last_order = C.order;
                    C.st.last_h = C.h;
                }
                __syncthreads();
                if (C.flag == 1) apply_change_D(C.order);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && g.cta == 0) {
            C.st.status = C.status;
            C.st.t_reached = C.t;
            C.st.n_out = C.k_out;
            A.stats[sysl] = C.st;
            if (A.prof) {
                PROF(7)
                for (int k = 0; k < 8; ++k) A.prof[(int64_t)sysl * 8 + k] = C.prof[k];
            }
        }
        group_sync(g);
    }
}
