// RHS / Jv over DENSE all-to-all pair blocks with the pair stream moved by the TMA engine
// (`cp.async.bulk` global -> shared, completion on mbarriers) instead of register loads.
//
// Why: the register version (stb_kernels.cuh) keeps 16 x 256 B per warp in flight, 64 KB per SM,
// and every outstanding line costs registers; on the 4.7 GB lattice it stops at 81 % of the
// measured copy peak.  Here every warp feeds its own row with 4 KB bulk copies into a ring of
// STAGES x 64 KB of shared memory (192 KB in flight per SM, no registers held by loads), and the
// 16 warps of the CTA walk the SAME partner columns of 16 different rows in lockstep, so the
// gathered populations z[k] are fetched from L2 once per CTA and stage and hit L1 for the other
// 15 rows.
//
//   CTA (persistent, one per SM) owns a contiguous range of rows of each ion type, cut into
//   near-equal tiles of RT <= 16 rows (every row costs the same in a dense block, so all CTAs
//   finish together).  The CTA's work is the sequence of ITEMS (tile, partner type U, chunk of KC
//   partner columns); item i lives in ring stage i % STAGES, slot [stage][w] belongs to warp w.
//       warp w, item i: wait full[stage] (all 16 rows of the item have landed);
//                       acc[s] += ring[stage][w][k] * z_U[col_s*n_U + k0 + k];
//                       lane 0: refill the slot with row w of item i + STAGES (expect_tx + one
//                       bulk copy, or a plain arrive when the tile has fewer rows) and prefetch
//                       1/16 of that item's z columns into L1.
//   The slot is private to the warp, so there is no "empty" barrier; the shared "full" barrier
//   (16 arrivals + bytes) keeps the warps within STAGES items of each other.  A single producer
//   warp is NOT used: `cp.async.bulk` takes its operands from uniform registers, a warp issues
//   divergent copies one lane at a time (ELECT / R2UR loop, ~90 ns per copy measured), and 15
//   copies per 1.4 us item made that loop the bottleneck.
//   After the last chunk of U: warp-shuffle row reduction -> field slots; after both U: the
//   warp-cooperative per-ion epilogue of stb_coop.cuh (block matvec + rate terms), lane = state.
//
// Algorithmic bytes are unchanged: 8 B per ordered pair + 16 N (RHS) / 24 N (Jv).  The per-lane
// summation order is that of stream_dense_t: results are bit-identical to k_rhs / k_jv.
// Reference semantics: odesolver._rate_eq / _rate_eq_pulse / _jac_rate_eq (odesolver.py:33-80).
#pragma once
#include "stb_kernels.cuh"

#ifndef STB_TMA_KC
#define STB_TMA_KC 512        // partner columns per stage and row (4 KB bulk copies)
#endif
#ifndef STB_TMA_STAGES
#define STB_TMA_STAGES 3
#endif
#define STB_TMA_ROWS 16       // rows per tile = warps of the CTA
#define STB_TMA_NT (STB_TMA_ROWS * 32)
#define STB_TMA_MAX_STAGES 16
// ring of `stages` x ROWS x `kc` doubles, then the full barriers; kc and stages are launch
// parameters (OpArgs.tma_kc / tma_stages; defaults above) so that they can be tuned per lattice
__host__ __device__ inline size_t tma_ring_bytes(int kc, int stages) {
    return (size_t)stages * STB_TMA_ROWS * kc * sizeof(double);
}
__host__ __device__ inline size_t tma_dyn_bytes(int kc, int stages) {
    return tma_ring_bytes(kc, stages) + STB_TMA_MAX_STAGES * sizeof(uint64_t);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared of this CTA; bytes % 16 == 0, both addresses 16 B aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct TmaRing {
    double* ring;             // [stages][ROWS][kc]
    uint64_t* full;           // [stages] 16 arrivals (one per warp) + the bytes of the item's rows
    int kc, stages;
    int diag;                 // measurement only: bit 0 = do not gather z (pure pair-stream ceiling),
                              // bit 1 = no L1 prefetch of z
    uint32_t stage, ph;       // ring position of the next item to consume (per thread, identical)
};

// carve the ring out of dynamic shared memory and initialise its barriers (all threads call)
__device__ __forceinline__ void tma_ring_init(TmaRing& R, void* smem, int kc, int stages, int diag) {
    R.kc = kc; R.stages = stages; R.diag = diag;
    R.ring = (double*)smem;
    R.full = (uint64_t*)((unsigned char*)smem + tma_ring_bytes(kc, stages));
    R.stage = 0; R.ph = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&R.full[s], STB_TMA_ROWS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// Position in the CTA's item sequence: types in the order (first_T, 1 - first_T), tiles of the
// CTA's rows [r_lo[T], r_hi[T]), partner types with field slots, chunks of kc columns.
struct TmaCursor {
    int Ti, tile, U, k0;      // loop indices
    int T, t0, RT;            // of the current item: row type, first row and rows of its tile
    bool done;
};

__device__ __forceinline__ void tma_seek(TmaCursor& C, const DevLattice& L, const int* r_lo, const int* r_hi,
                                         int first_T) {
    for (;;) {
        if (C.Ti >= 2) { C.done = true; return; }
        const int T = C.Ti == 0 ? first_T : 1 - first_T;
        const int rows = max(r_hi[T] - r_lo[T], 0);
        const int ntile = (rows + STB_TMA_ROWS - 1) / STB_TMA_ROWS;
        if (C.tile >= ntile) { ++C.Ti; C.tile = 0; C.U = 0; C.k0 = 0; continue; }
        if (C.U >= 2) { ++C.tile; C.U = 0; C.k0 = 0; continue; }
        if (L.blk[T][C.U].n_slots == 0 || L.n[C.U] == 0 || C.k0 >= L.ld[C.U]) { ++C.U; C.k0 = 0; continue; }
        C.T = T;
        C.t0 = r_lo[T] + (int)((int64_t)C.tile * rows / ntile);
        C.RT = r_lo[T] + (int)((int64_t)(C.tile + 1) * rows / ntile) - C.t0;
        return;
    }
}

// Warp w's share of issuing item C into ring stage `stage`: its row's bulk copy (or a bare
// arrival) and 1/16 of the L1 prefetch of the item's gathered columns.
__device__ __forceinline__ void tma_issue(const TmaRing& R, const DevLattice& L, const TmaCursor& C,
                                          const double* x, uint32_t stage, int w, int lane) {
    const PairBlock& B = L.blk[C.T][C.U];
    const int ld = L.ld[C.U], nU = L.n[C.U];
    const int kc = min(R.kc, ld - C.k0);
    if (lane == 0) {
        if (w < C.RT) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&R.full[stage], (uint32_t)(kc * (int)sizeof(double)));
            bulk_g2s(R.ring + ((size_t)stage * STB_TMA_ROWS + w) * R.kc,
                     B.g[0] + (int64_t)(C.t0 + w) * ld + C.k0, (uint32_t)(kc * (int)sizeof(double)),
                     &R.full[stage]);
        } else {
            mbar_arrive(&R.full[stage]);
        }
    }
    if (!(R.diag & 2) && lane < 2 * B.n_slots) {       // 128 B lines 2w, 2w + 1 of every slot's columns
        const int k = C.k0 + (2 * w + (lane & 1)) * 16;
        if (k < C.k0 + kc)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(x + L.off[C.U] + B.slot_col[lane >> 1] * nU + min(k, nU - 1)));
    }
}

// What the issue side needs to run `stages` items ahead of the consumer loop.
struct TmaFeed {
    TmaCursor C;
    const int* r_lo;
    const int* r_hi;
    int first_T;
    const double* x;
};

__device__ __forceinline__ void tma_feed_next(const TmaRing& R, const DevLattice& L, TmaFeed& Fd, uint32_t stage,
                                              int w, int lane) {
    if (Fd.C.done) return;
    tma_issue(R, L, Fd.C, Fd.x, stage, w, lane);
    Fd.C.k0 += R.kc;
    tma_seek(Fd.C, L, Fd.r_lo, Fd.r_hi, Fd.first_T);
}

// Consumer side of one (tile, partner type) block: accumulates the NS field slots of row `w` of
// the tile over all column chunks, refilling the warp's ring slot after every chunk.
template <int NS>
__device__ __forceinline__ void tma_consume_block(TmaRing& R, const DevLattice& L, TmaFeed& Fd, const PairBlock& B,
                                                  int nU, int ld, const double* xU, int w, bool has_row,
                                                  int lane, double (&acc)[STB_MAXC]) {
    int sc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) sc[s] = B.slot_col[s] * nU;
    const int kmax = nU - 1;
    constexpr int JB = NS <= 2 ? 16 : (NS <= 4 ? 8 : 4);      // column lines whose z loads are in flight together
    for (int k0 = 0; k0 < ld; k0 += R.kc) {
        const int kc = min(R.kc, ld - k0);
        const uint32_t stage = R.stage, ph = R.ph;
        const double* grow = R.ring + ((size_t)stage * STB_TMA_ROWS + w) * R.kc + lane;
        const int nj = kc >> 5;
        if (has_row && kc == STB_TMA_KC && k0 + kc <= nU) {
            // full chunk of the default width: compile-time trip counts, no clamps
            const double* zk = xU + k0 + lane;
#pragma unroll
            for (int j0 = 0; j0 < STB_TMA_KC / 32; j0 += JB) {
                double z[JB][NS];
#pragma unroll
                for (int u = 0; u < JB; ++u) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) z[u][s] = (R.diag & 1) ? 1.0 : zk[sc[s] + 32 * (j0 + u)];
                }
                if (j0 == 0) mbar_wait(&R.full[stage], ph);
#pragma unroll
                for (int u = 0; u < JB; ++u) {
                    const double g = grow[32 * (j0 + u)];
#pragma unroll
                    for (int s = 0; s < NS; ++s) acc[s] = fma(g, z[u][s], acc[s]);
                }
            }
        } else if (has_row) {
            for (int j0 = 0; j0 < nj; j0 += JB) {
                double z[JB][NS];
#pragma unroll
                for (int u = 0; u < JB; ++u) {
                    const int k = min(k0 + lane + 32 * min(j0 + u, nj - 1), kmax);
#pragma unroll
                    for (int s = 0; s < NS; ++s) z[u][s] = (R.diag & 1) ? 1.0 : xU[sc[s] + k];
                }
                if (j0 == 0) mbar_wait(&R.full[stage], ph);
#pragma unroll
                for (int u = 0; u < JB; ++u) {
                    const double g = j0 + u < nj ? grow[32 * (j0 + u)] : 0.0;
#pragma unroll
                    for (int s = 0; s < NS; ++s) acc[s] = fma(g, z[u][s], acc[s]);
                }
            }
        } else {
            mbar_wait(&R.full[stage], ph);
        }
        __syncwarp();
        tma_feed_next(R, L, Fd, stage, w, lane);      // item + stages goes into the slot just read
        if (++R.stage == (uint32_t)R.stages) { R.stage = 0; R.ph ^= 1u; }
    }
}

// One pass over the CTA's rows: for every row, `epi(T, q, true, 32, lane, fl)` is called by the
// row's warp with lane k holding field slot k (times fscale) of the row, gathered from `x`.
// All 16 warps of the CTA must call this together (same arguments).
template <class Epi>
__device__ __forceinline__ void tma_rows_pass(TmaRing& R, const DevLattice& L, const int* r_lo, const int* r_hi,
                                              int first_T, const double* x, double fscale, int w, int lane,
                                              Epi&& epi) {
    TmaFeed Fd;
    Fd.C.Ti = 0; Fd.C.tile = 0; Fd.C.U = 0; Fd.C.k0 = 0; Fd.C.T = 0; Fd.C.t0 = 0; Fd.C.RT = 0; Fd.C.done = false;
    Fd.r_lo = r_lo; Fd.r_hi = r_hi; Fd.first_T = first_T; Fd.x = x;
    tma_seek(Fd.C, L, r_lo, r_hi, first_T);
    {   // prologue: the first `stages` items (the ring is empty between passes)
        uint32_t st = R.stage;
        for (int i = 0; i < R.stages; ++i) {
            tma_feed_next(R, L, Fd, st, w, lane);
            if (++st == (uint32_t)R.stages) st = 0;
        }
    }
    for (int Ti = 0; Ti < 2; ++Ti) {
        const int T = Ti == 0 ? first_T : 1 - first_T;
        const int rows = r_hi[T] - r_lo[T];
        if (rows <= 0) continue;
        const int ntile = (rows + STB_TMA_ROWS - 1) / STB_TMA_ROWS;
        for (int tile = 0; tile < ntile; ++tile) {
            const int t0 = r_lo[T] + (int)((int64_t)tile * rows / ntile);
            const int RT = r_lo[T] + (int)((int64_t)(tile + 1) * rows / ntile) - t0;
            const bool has_row = w < RT;
            int slot0 = 0;
            double fl = 0.0;
            for (int U = 0; U < 2; ++U) {
                const PairBlock& B = L.blk[T][U];
                const int ns = B.n_slots, nU = L.n[U];
                if (ns == 0) continue;
                if (nU > 0) {
                    double acc[STB_MAXC];
#pragma unroll
                    for (int s = 0; s < STB_MAXC; ++s) acc[s] = 0.0;
                    const double* xU = x + L.off[U];
                    switch (ns) {
                        case 1: tma_consume_block<1>(R, L, Fd, B, nU, L.ld[U], xU, w, has_row, lane, acc); break;
                        case 2: tma_consume_block<2>(R, L, Fd, B, nU, L.ld[U], xU, w, has_row, lane, acc); break;
                        case 3: tma_consume_block<3>(R, L, Fd, B, nU, L.ld[U], xU, w, has_row, lane, acc); break;
                        case 4: tma_consume_block<4>(R, L, Fd, B, nU, L.ld[U], xU, w, has_row, lane, acc); break;
                        case 5: tma_consume_block<5>(R, L, Fd, B, nU, L.ld[U], xU, w, has_row, lane, acc); break;
                        default: tma_consume_block<6>(R, L, Fd, B, nU, L.ld[U], xU, w, has_row, lane, acc); break;
                    }
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int s = 0; s < STB_MAXC; ++s)
                            if (s < ns) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
                    }
#pragma unroll
                    for (int s = 0; s < STB_MAXC; ++s)
                        if (s < ns && lane == slot0 + s) fl = acc[s] * fscale;
                }
                slot0 += ns;
            }
            if (has_row) epi(T, t0 + w, true, 32, lane, fl);
        }
    }
}

template <bool JV>
__global__ void __launch_bounds__(STB_TMA_NT, 1) k_op_tma(OpArgs A) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ DevLattice L;
    __shared__ SysParams P;
    for (int i = threadIdx.x; i < (int)(sizeof(DevLattice) / 4); i += blockDim.x)
        ((int*)&L)[i] = ((const int*)A.L)[i];
    TmaRing R;
    tma_ring_init(R, tma_smem, A.tma_kc, A.tma_stages, A.tma_diag);     // includes __syncthreads
    const int b = blockIdx.y;
    load_sys_params(L, A.blocks, A.strengths, A.first + b, A.pulse, P);
    __syncthreads();
    const double* y = A.y + (int64_t)b * L.N;
    const double* x = JV ? A.v + (int64_t)b * L.N : y;       // the vector whose fields are streamed
    const double* Fy = JV ? A.Fy + (int64_t)b * L.NF : nullptr;
    double* out = A.out + (int64_t)b * L.N;
    double* F = (!JV && A.F) ? A.F + (int64_t)b * L.NF : nullptr;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int r_lo[2], r_hi[2];
    for (int T = 0; T < 2; ++T) {
        r_lo[T] = (int)((int64_t)blockIdx.x * L.n[T] / gridDim.x);
        r_hi[T] = (int)((int64_t)(blockIdx.x + 1) * L.n[T] / gridDim.x);
    }
    // warp-cooperative epilogue: lane j <-> state j of ion q, lane k <-> field slot k
    tma_rows_pass(R, L, r_lo, r_hi, 0, x, 1.0, w, lane, [&](int T, int q, bool, int, int, double fl) {
        const int s = L.s[T], nT = L.n[T];
        const int e = L.off[T] + (lane < s ? lane : 0) * nT + q;
        if (!JV) {
            const double yj = lane < s ? y[e] : 0.0;
            const double fj = coop_rhs(L, P, T, 32, lane, yj, fl);
            if (lane < s) out[e] = fj;
            if (F && lane < L.nf[T]) F[L.foff[T] + lane * nT + q] = fl;
        } else {
            const double yj = lane < s ? y[e] : 0.0, vj = lane < s ? x[e] : 0.0;
            const double fyl = lane < L.nf[T] ? Fy[L.foff[T] + lane * nT + q] : 0.0;
            const double oj = coop_jv(L, P, T, 32, lane, yj, vj, fyl, fl);
            if (lane < s) out[e] = oj;
        }
    });
}
