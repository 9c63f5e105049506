// Pair-field streaming and per-ion epilogues shared by the standalone RHS / Jv kernels and the
// persistent integrator.
//
// For row ion q (type T) and partner type U the kernel streams the row's sliced-ELL entries
// (partner index c, g = d^-mult) and accumulates, per field slot (state column sc, mult m):
//       F[slot] = sum_k  g_m(q,k) * z_U[sc * n_U + c_k]
// This is the only part that touches the pair data: 4 + 8*n_mult bytes per stored pair.
// Everything else (block matvec with the decay/absorption blocks, the ET rate terms, the Jv
// combination, the block-Jacobi solve) is per-ion work done by the row's lead lane from the
// reduced fields - the reference's three scipy calls of odesolver.py:37-43 fused into one pass.
#pragma once
#include "stb_common.cuh"

__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) { return __ldcs(p); }

template <int NS, int NM>
__device__ __forceinline__ void stream_block_t(const PairBlock& B, int nU, const double* zU,
                                               int64_t l0, int64_t l1, int lane,
                                               double (&acc)[STB_MAXC]) {
    int sc[NS], sm[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        sc[s] = B.slot_col[s] * nU;
        sm[s] = B.slot_mult[s];
    }
    const uint32_t* col = B.col;
    const double* g0p = B.g[0];
    const double* g1p = (NM > 1 && B.g[1]) ? B.g[1] : B.g[0];
    const double* g2p = (NM > 2 && B.g[2]) ? B.g[2] : B.g[0];
    int64_t ln = l0;
    constexpr int UN = 4;
    for (; ln + UN <= l1; ln += UN) {
        uint32_t c[UN];
        double g0[UN], g1[UN], g2[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            int64_t idx = (ln + u) * 32 + lane;
            c[u] = ld_stream(col + idx);
            g0[u] = ld_stream(g0p + idx);
            if (NM > 1) g1[u] = ld_stream(g1p + idx);
            if (NM > 2) g2[u] = ld_stream(g2p + idx);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                double gg = g0[u];
                if (NM > 1) gg = sm[s] == 0 ? g0[u] : (sm[s] == 1 || NM == 2 ? g1[u] : g2[u]);
                acc[s] = fma(gg, zU[sc[s] + c[u]], acc[s]);
            }
        }
    }
    for (; ln < l1; ++ln) {
        int64_t idx = ln * 32 + lane;
        uint32_t c = ld_stream(col + idx);
        double g0 = ld_stream(g0p + idx), g1 = 0, g2 = 0;
        if (NM > 1) g1 = ld_stream(g1p + idx);
        if (NM > 2) g2 = ld_stream(g2p + idx);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double gg = g0;
            if (NM > 1) gg = sm[s] == 0 ? g0 : (sm[s] == 1 || NM == 2 ? g1 : g2);
            acc[s] = fma(gg, zU[sc[s] + c], acc[s]);
        }
    }
}

template <int NM>
__device__ __forceinline__ void stream_block_m(const PairBlock& B, int nU, const double* zU,
                                               int64_t l0, int64_t l1, int lane,
                                               double (&acc)[STB_MAXC]) {
    switch (B.n_slots) {
        case 1: stream_block_t<1, NM>(B, nU, zU, l0, l1, lane, acc); break;
        case 2: stream_block_t<2, NM>(B, nU, zU, l0, l1, lane, acc); break;
        case 3: stream_block_t<3, NM>(B, nU, zU, l0, l1, lane, acc); break;
        case 4: stream_block_t<4, NM>(B, nU, zU, l0, l1, lane, acc); break;
        case 5: stream_block_t<5, NM>(B, nU, zU, l0, l1, lane, acc); break;
        case 6: stream_block_t<6, NM>(B, nU, zU, l0, l1, lane, acc); break;
        default: break;
    }
}

// Fields of the rows of slice `slice` of type T gathered from the plane vector z, multiplied by
// `scale`.  On return every lane of a row holds the row's reduced sums in Fq[0..nf[T]):
// partner-S slots first, then partner-A slots.  One (non-inlined) instance serves the RHS, Jv
// and integrator kernels.
__device__ __noinline__ void slice_fields(const DevLattice& L, int T, int slice, const double* z,
                                          int lane, double scale, double* Fq) {
    double accS[STB_MAXC], accA[STB_MAXC];
#pragma unroll
    for (int s = 0; s < STB_MAXC; ++s) accS[s] = accA[s] = 0.0;
#pragma unroll
    for (int U = 0; U < 2; ++U) {
        const PairBlock& B = L.blk[T][U];
        if (B.n_slots == 0) continue;
        int64_t l0 = B.line_ptr[slice], l1 = B.line_ptr[slice + 1];
        const double* zU = z + L.off[U];
        if (U == 0) {
            if (L.n_mult <= 1) stream_block_m<1>(B, L.n[U], zU, l0, l1, lane, accS);
            else stream_block_m<3>(B, L.n[U], zU, l0, l1, lane, accS);
        } else {
            if (L.n_mult <= 1) stream_block_m<1>(B, L.n[U], zU, l0, l1, lane, accA);
            else stream_block_m<3>(B, L.n[U], zU, l0, l1, lane, accA);
        }
    }
    const int n0 = L.blk[T][0].n_slots, n1 = L.blk[T][1].n_slots;
    for (int o = L.lpr >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int s = 0; s < STB_MAXC; ++s) {
            if (s < n0) accS[s] += __shfl_xor_sync(0xffffffffu, accS[s], o);
            if (s < n1) accA[s] += __shfl_xor_sync(0xffffffffu, accA[s], o);
        }
    }
#pragma unroll
    for (int s = 0; s < STB_MAXC; ++s) {
        if (s < n0) Fq[s] = accS[s] * scale;
    }
#pragma unroll
    for (int s = 0; s < STB_MAXC; ++s) {
        if (s < n1) Fq[n0 + s] = accA[s] * scale;
    }
}

__device__ __forceinline__ void load_row(const DevLattice& L, int T, int q, const double* z,
                                         double* zq) {
    const double* p = z + L.off[T] + q;
    int nT = L.n[T];
    for (int a = 0; a < L.s[T]; ++a) zq[a] = p[(int64_t)a * nT];
}

__device__ __forceinline__ void store_row(const DevLattice& L, int T, int q, double* z,
                                          const double* zq) {
    double* p = z + L.off[T] + q;
    int nT = L.n[T];
    for (int a = 0; a < L.s[T]; ++a) p[(int64_t)a * nT] = zq[a];
}

// out = Lin_T yq + sum_t C_t yq[a_t] Fq[slot_t] (e_b - e_a)          (odesolver.py:33-43,65-71)
__device__ __forceinline__ void row_rhs(const DevLattice& L, const SysParams& P, int T,
                                        const double* yq, const double* Fq, double* out) {
    int s = L.s[T];
    const double* lin = P.lin[T];
    for (int a = 0; a < s; ++a) {
        double acc = 0.0;
        for (int b = 0; b < s; ++b) acc = fma(lin[a * s + b], yq[b], acc);
        out[a] = acc;
    }
    const TermTable& tt = L.terms[T];
    for (int t = 0; t < tt.n_terms; ++t) {
        double rate = P.C[tt.proc[t]] * yq[tt.a[t]] * Fq[tt.fslot[t]];
        out[tt.a[t]] -= rate;
        out[tt.b[t]] += rate;
    }
}

// out = Lin_T vq + sum_t C_t (vq[a] Fy[slot] + yq[a] Fv[slot]) (e_b - e_a)   (odesolver.py:46-62)
__device__ __forceinline__ void row_jv(const DevLattice& L, const SysParams& P, int T,
                                       const double* yq, const double* vq, const double* Fy,
                                       const double* Fv, double* out) {
    int s = L.s[T];
    const double* lin = P.lin[T];
    for (int a = 0; a < s; ++a) {
        double acc = 0.0;
        for (int b = 0; b < s; ++b) acc = fma(lin[a * s + b], vq[b], acc);
        out[a] = acc;
    }
    const TermTable& tt = L.terms[T];
    for (int t = 0; t < tt.n_terms; ++t) {
        int a = tt.a[t], f = tt.fslot[t];
        double rate = P.C[tt.proc[t]] * fma(vq[a], Fy[f], yq[a] * Fv[f]);
        out[a] -= rate;
        out[tt.b[t]] += rate;
    }
}

// Minv = (I - c * Jblock)^-1 with Jblock = Lin_T + sum_t C_t Fy[slot] (e_b - e_a) e_a^T
// (the part of the Jacobian whose row and column belong to ion q).  Gauss-Jordan with partial
// pivoting on an s x s block held in local memory.
__device__ __forceinline__ void row_block_inverse(const DevLattice& L, const SysParams& P, int T,
                                                  const double* Fy, double c, double* A /*s*s*/,
                                                  double* Ainv /*s*s*/) {
    int s = L.s[T];
    const double* lin = P.lin[T];
    for (int i = 0; i < s * s; ++i) A[i] = -c * lin[i];
    const TermTable& tt = L.terms[T];
    for (int t = 0; t < tt.n_terms; ++t) {
        int a = tt.a[t], b = tt.b[t];
        double k = c * P.C[tt.proc[t]] * Fy[tt.fslot[t]];
        A[a * s + a] += k;
        A[b * s + a] -= k;
    }
    for (int i = 0; i < s; ++i) A[i * s + i] += 1.0;
    for (int i = 0; i < s * s; ++i) Ainv[i] = 0.0;
    for (int i = 0; i < s; ++i) Ainv[i * s + i] = 1.0;
    for (int col = 0; col < s; ++col) {
        int piv = col;
        double best = fabs(A[col * s + col]);
        for (int r = col + 1; r < s; ++r) {
            double v = fabs(A[r * s + col]);
            if (v > best) { best = v; piv = r; }
        }
        if (piv != col) {
            for (int k = 0; k < s; ++k) {
                double t0 = A[col * s + k]; A[col * s + k] = A[piv * s + k]; A[piv * s + k] = t0;
                double t1 = Ainv[col * s + k]; Ainv[col * s + k] = Ainv[piv * s + k]; Ainv[piv * s + k] = t1;
            }
        }
        double d = A[col * s + col];
        double inv = d != 0.0 ? 1.0 / d : 0.0;
        for (int k = 0; k < s; ++k) { A[col * s + k] *= inv; Ainv[col * s + k] *= inv; }
        for (int r = 0; r < s; ++r) {
            if (r == col) continue;
            double m = A[r * s + col];
            if (m == 0.0) continue;
            for (int k = 0; k < s; ++k) {
                A[r * s + k] = fma(-m, A[col * s + k], A[r * s + k]);
                Ainv[r * s + k] = fma(-m, Ainv[col * s + k], Ainv[r * s + k]);
            }
        }
    }
}

__device__ __forceinline__ void load_sys_params(const DevLattice& L, const double* blocks,
                                                const double* strengths, int sys, int pulse,
                                                SysParams& P) {
    // blocks: [n_sys][2 (decay, abs)][sS*sS + sA*sA]
    int per = L.s[0] * L.s[0] + L.s[1] * L.s[1];
    const double* dec = blocks + (int64_t)sys * 2 * per;
    const double* ab = dec + per;
    for (int i = threadIdx.x; i < per; i += blockDim.x) {
        int T = i < L.s[0] * L.s[0] ? 0 : 1;
        int k = T == 0 ? i : i - L.s[0] * L.s[0];
        P.lin[T][k] = dec[i] + (pulse ? ab[i] : 0.0);
    }
    for (int i = threadIdx.x; i < L.n_proc; i += blockDim.x) {
        double c = strengths[(int64_t)sys * L.n_proc + i];
        P.C[i] = stb_active_strength(c) ? c : 0.0;
    }
}
