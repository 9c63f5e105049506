// RHS / Jv over DENSE all-to-all pair blocks with the pair stream moved by the TMA engine
// (`cp.async.bulk` global -> shared, completion on mbarriers) instead of register loads.
//
// Why: the register version (stb_kernels.cuh) keeps 16 x 256 B per warp in flight, 64 KB per SM,
// and every outstanding line costs registers; on the 4.7 GB lattice it stops at 81 % of the
// measured copy peak.  Here one elected lane per row issues 4 KB bulk copies into a ring of
// STAGES x 60 KB of shared memory (180 KB in flight per SM, no registers held by loads), and the
// 15 consumer warps of the CTA walk the SAME partner columns of 15 different rows in lockstep, so
// the gathered populations z[k] are fetched from L2 once per CTA and stage and hit L1 for the
// other 14 rows.
//
//   CTA (persistent, one per SM) owns a contiguous range of rows of each ion type, cut into
//   near-equal tiles of RT <= 15 rows (every row costs the same in a dense block, so all CTAs
//   finish together).  For a tile, for each partner type U, for each chunk of KC partner columns:
//       producer warp : wait empty[stage]; expect RT*kc*8 bytes on full[stage];
//                       lane r < RT copies g[(t0+r)*ld + k0 .. +kc) -> ring[stage][r][:]
//       consumer warp w: wait full[stage]; acc[s] += ring[stage][w][k] * z_U[col_s*n_U + k0 + k];
//                       arrive on empty[stage]
//   after the last chunk of U: warp-shuffle row reduction -> field slots; after both U: the
//   warp-cooperative per-ion epilogue of stb_coop.cuh (block matvec + rate terms), lane = state.
//
// Algorithmic bytes are unchanged: 8 B per ordered pair + 16 N (RHS) / 24 N (Jv).
// Reference semantics: odesolver._rate_eq / _rate_eq_pulse / _jac_rate_eq (odesolver.py:33-80).
#pragma once
#include "stb_kernels.cuh"

#ifndef STB_TMA_KC
#define STB_TMA_KC 512        // partner columns per stage and row (4 KB bulk copies)
#endif
#ifndef STB_TMA_STAGES
#define STB_TMA_STAGES 3
#endif
#define STB_TMA_ROWS 15       // rows per tile = consumer warps (15 + the producer warp = 512 threads, 128 registers each)
#define STB_TMA_NT ((STB_TMA_ROWS + 1) * 32)
#define STB_TMA_RING_BYTES ((size_t)STB_TMA_STAGES * STB_TMA_ROWS * STB_TMA_KC * sizeof(double))
#define STB_TMA_DYN_BYTES (STB_TMA_RING_BYTES + 2 * STB_TMA_STAGES * sizeof(uint64_t))

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
    double* ring;             // [STAGES][ROWS][KC]
    uint64_t* full;           // [STAGES] producer -> consumers (transaction bytes)
    uint64_t* empty;          // [STAGES] consumers -> producer (one arrival per consumer warp)
};

// Consumer side of one (tile, partner type) block: accumulates the NS field slots of row
// `w` of the tile over all column chunks.  `it` = running stage-item counter (identical sequence
// in the producer).  Every consumer warp takes part in the barrier protocol, rows or not.
template <int NS>
__device__ __forceinline__ void tma_consume_block(const TmaRing& R, const PairBlock& B, int nU, int ld,
                                                  const double* xU, int w, bool has_row, int lane,
                                                  uint32_t& it, double (&acc)[STB_MAXC]) {
    int sc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) sc[s] = B.slot_col[s] * nU;
    const int kmax = nU - 1;
    constexpr int JB = NS <= 2 ? 16 : (NS <= 4 ? 8 : 4);      // column lines whose z loads are in flight together
    for (int k0 = 0; k0 < ld; k0 += STB_TMA_KC, ++it) {
        const int kc = min(STB_TMA_KC, ld - k0);
        const int stage = it % STB_TMA_STAGES;
        const uint32_t ph = (it / STB_TMA_STAGES) & 1u;
        const double* grow = R.ring + ((size_t)stage * STB_TMA_ROWS + w) * STB_TMA_KC + lane;
        const int nj = kc >> 5;
        if (has_row) {
            for (int j0 = 0; j0 < nj; j0 += JB) {
                double z[JB][NS];
#pragma unroll
                for (int u = 0; u < JB; ++u) {
                    const int k = min(k0 + lane + 32 * min(j0 + u, nj - 1), kmax);
#pragma unroll
                    for (int s = 0; s < NS; ++s) z[u][s] = xU[sc[s] + k];
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
        if (lane == 0) mbar_arrive(&R.empty[stage]);
    }
}

// Producer side of the same block: one bulk copy per row and chunk.
__device__ __forceinline__ void tma_produce_block(const TmaRing& R, const PairBlock& B, int ld, int t0,
                                                  int RT, int lane, uint32_t& it) {
    for (int k0 = 0; k0 < ld; k0 += STB_TMA_KC, ++it) {
        const int kc = min(STB_TMA_KC, ld - k0);
        const int stage = it % STB_TMA_STAGES;
        const uint32_t ph = (it / STB_TMA_STAGES) & 1u;
        mbar_wait(&R.empty[stage], ph ^ 1u);
        if (lane == 0) mbar_expect_tx(&R.full[stage], (uint32_t)(RT * kc * (int)sizeof(double)));
        __syncwarp();
        if (lane < RT)
            bulk_g2s(R.ring + ((size_t)stage * STB_TMA_ROWS + lane) * STB_TMA_KC,
                     B.g[0] + (int64_t)(t0 + lane) * ld + k0, (uint32_t)(kc * (int)sizeof(double)),
                     &R.full[stage]);
    }
}

template <bool JV>
__global__ void __launch_bounds__(STB_TMA_NT, 1) k_op_tma(OpArgs A) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ DevLattice L;
    __shared__ SysParams P;
    TmaRing R;
    R.ring = (double*)tma_smem;
    R.full = (uint64_t*)(tma_smem + STB_TMA_RING_BYTES);
    R.empty = R.full + STB_TMA_STAGES;
    for (int i = threadIdx.x; i < (int)(sizeof(DevLattice) / 4); i += blockDim.x)
        ((int*)&L)[i] = ((const int*)A.L)[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STB_TMA_STAGES; ++s) {
            mbar_init(&R.full[s], 1);
            mbar_init(&R.empty[s], STB_TMA_ROWS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int b = blockIdx.y;
    load_sys_params(L, A.blocks, A.strengths, A.first + b, A.pulse, P);
    __syncthreads();
    const double* y = A.y + (int64_t)b * L.N;
    const double* x = JV ? A.v + (int64_t)b * L.N : y;       // the vector whose fields are streamed
    const double* Fy = JV ? A.Fy + (int64_t)b * L.NF : nullptr;
    double* out = A.out + (int64_t)b * L.N;
    double* F = (!JV && A.F) ? A.F + (int64_t)b * L.NF : nullptr;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool producer = w == STB_TMA_ROWS;
    uint32_t it = 0;
    for (int T = 0; T < 2; ++T) {
        const int nT = L.n[T];
        const int r0 = (int)((int64_t)blockIdx.x * nT / gridDim.x);
        const int r1 = (int)((int64_t)(blockIdx.x + 1) * nT / gridDim.x);
        const int rows = r1 - r0;
        if (rows <= 0) continue;
        const int ntile = (rows + STB_TMA_ROWS - 1) / STB_TMA_ROWS;
        for (int tile = 0; tile < ntile; ++tile) {
            const int t0 = r0 + (int)((int64_t)tile * rows / ntile);
            const int RT = r0 + (int)((int64_t)(tile + 1) * rows / ntile) - t0;
            if (producer) {
                for (int U = 0; U < 2; ++U) {
                    const PairBlock& B = L.blk[T][U];
                    if (B.n_slots == 0 || L.n[U] == 0) continue;
                    tma_produce_block(R, B, L.ld[U], t0, RT, lane, it);
                }
                continue;
            }
            const bool has_row = w < RT;
            const int q = has_row ? t0 + w : t0;
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
                        case 1: tma_consume_block<1>(R, B, nU, L.ld[U], xU, w, has_row, lane, it, acc); break;
                        case 2: tma_consume_block<2>(R, B, nU, L.ld[U], xU, w, has_row, lane, it, acc); break;
                        case 3: tma_consume_block<3>(R, B, nU, L.ld[U], xU, w, has_row, lane, it, acc); break;
                        case 4: tma_consume_block<4>(R, B, nU, L.ld[U], xU, w, has_row, lane, it, acc); break;
                        case 5: tma_consume_block<5>(R, B, nU, L.ld[U], xU, w, has_row, lane, it, acc); break;
                        default: tma_consume_block<6>(R, B, nU, L.ld[U], xU, w, has_row, lane, it, acc); break;
                    }
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int s = 0; s < STB_MAXC; ++s)
                            if (s < ns) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
                    }
#pragma unroll
                    for (int s = 0; s < STB_MAXC; ++s)
                        if (s < ns && lane == slot0 + s) fl = acc[s];
                }
                slot0 += ns;
            }
            if (!has_row) continue;
            // warp-cooperative epilogue: lane j <-> state j of ion q, lane k <-> field slot k
            const int s = L.s[T];
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
        }
    }
}
