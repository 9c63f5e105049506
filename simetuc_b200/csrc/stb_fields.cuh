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

// Pair data is read exactly once per pass: bypass L1 (it is needed for the gathered state
// vector) and keep the default L2 policy (small lattices stay L2-resident across passes).
#ifndef STB_LD_V
#define STB_LD_V 2
#endif
#if STB_LD_V == 1
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) { return __ldcs(p); }
#else
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {
    uint32_t v;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
#endif

#ifndef STB_STREAM_V
#define STB_STREAM_V 2
#endif
#ifndef STB_UN_KERNEL
#define STB_UN_KERNEL 16  // lines in flight per lane, standalone RHS / Jv kernels (HBM-bound)
#endif
#ifndef STB_UN_INT
#define STB_UN_INT 16     // ... inside the persistent integrator (measured on the 4.7 GB lattice:
                          // 4 lines in flight 1.39 ms per pass, 16 lines 1.10 ms; no spills either way)
#endif

// One multipolarity (the common case): NS field slots, accumulators in registers.
template <class T_>
__device__ __forceinline__ T_ ld_pair(const T_* p, bool nc) { return nc ? ld_stream(p) : *p; }

// NC = true: pair data in global memory, read through the non-allocating path;
// NC = false: generic loads (the pair data of the CTA's rows is resident in shared memory).
template <int NS, int STB_UN, bool NC>
__device__ __forceinline__ void stream_block_t(const PairBlock& B, int nU, const double* zU,
                                               int64_t l0, int64_t l1, int lane,
                                               double (&acc)[STB_MAXC]) {
    int sc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) sc[s] = B.slot_col[s] * nU;
    const uint32_t* col = B.col + lane;
    const double* g0p = B.g[0] + lane;
    int64_t ln = l0;
    // full blocks: STB_UN lines in flight per lane, no predication (the HBM-bound regime)
    for (; ln + STB_UN <= l1; ln += STB_UN) {
        uint32_t c[STB_UN];
        double g[STB_UN];
#pragma unroll
        for (int u = 0; u < STB_UN; ++u) {
            c[u] = ld_pair(col + (ln + u) * 32, NC);
            g[u] = ld_pair(g0p + (ln + u) * 32, NC);
        }
#pragma unroll
        for (int u = 0; u < STB_UN; ++u) {
#pragma unroll
            for (int s = 0; s < NS; ++s) acc[s] = fma(g[u], zU[sc[s] + c[u]], acc[s]);
        }
    }
    // remainder in blocks of 4 clamped lines so that its loads overlap (short rows are
    // latency-bound); lines past the end are masked when consumed
    for (; ln < l1; ln += 4) {
        uint32_t c[4];
        double g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t lc = ln + u < l1 ? ln + u : l1 - 1;
            c[u] = ld_pair(col + lc * 32, NC);
            g[u] = ld_pair(g0p + lc * 32, NC);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double gu = ln + u < l1 ? g[u] : 0.0;
#pragma unroll
            for (int s = 0; s < NS; ++s) acc[s] = fma(gu, zU[sc[s] + c[u]], acc[s]);
        }
    }
}

// Several multipolarities: generic (runtime slot count), one line at a time.
__device__ __forceinline__ void stream_block_multi(const PairBlock& B, int nU, const double* zU,
                                                   int64_t l0, int64_t l1, int lane,
                                                   double (&acc)[STB_MAXC]) {
    const int ns = B.n_slots;
    for (int64_t ln = l0; ln < l1; ++ln) {
        const int64_t idx = ln * 32 + lane;
        const uint32_t c = ld_stream(B.col + idx);
        double g[STB_MAXM];
#pragma unroll
        for (int m = 0; m < STB_MAXM; ++m) g[m] = B.g[m] ? ld_stream(B.g[m] + idx) : 0.0;
#pragma unroll
        for (int s = 0; s < STB_MAXC; ++s) {
            if (s < ns) {
                const int m = B.slot_mult[s];
                const double gg = m == 0 ? g[0] : (m == 1 ? g[1] : g[2]);
                acc[s] = fma(gg, zU[B.slot_col[s] * nU + c], acc[s]);
            }
        }
    }
}

// ---- dense all-to-all blocks: g[q * ld + k], partner k = column index, no index stream ------
// A row is read by its `lpr` lanes, lane gl takes k = gl, gl + lpr, ...; the gathered
// populations z[sc + k] are contiguous across the lanes.  Padding (k >= nU) holds g = 0 and the
// population index is clamped.
template <int NS, int STB_UN>
__device__ __forceinline__ void stream_dense_t(const PairBlock& B, int nU, int ld, int lpr,
                                               const double* zU, int q, int gl,
                                               double (&acc)[STB_MAXC]) {
    int sc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) sc[s] = B.slot_col[s] * nU;
    const double* gp = B.g[0] + (int64_t)q * ld + gl;
    const int nl = ld / lpr, kmax = nU - 1;
    int i = 0;
    for (; i + STB_UN <= nl; i += STB_UN) {
        double g[STB_UN];
#pragma unroll
        for (int u = 0; u < STB_UN; ++u) g[u] = ld_stream(gp + (i + u) * lpr);
#pragma unroll
        for (int u = 0; u < STB_UN; ++u) {
            const int k = min(gl + (i + u) * lpr, kmax);
#pragma unroll
            for (int s = 0; s < NS; ++s) acc[s] = fma(g[u], zU[sc[s] + k], acc[s]);
        }
    }
    for (; i < nl; i += 4) {
        double g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) g[u] = ld_stream(gp + (i + u < nl ? i + u : nl - 1) * lpr);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double gu = i + u < nl ? g[u] : 0.0;
            const int k = min(gl + (i + u) * lpr, kmax);
#pragma unroll
            for (int s = 0; s < NS; ++s) acc[s] = fma(gu, zU[sc[s] + k], acc[s]);
        }
    }
}

__device__ __forceinline__ void stream_dense_multi(const PairBlock& B, int nU, int ld, int lpr,
                                                   const double* zU, int q, int gl,
                                                   double (&acc)[STB_MAXC]) {
    const int ns = B.n_slots, kmax = nU - 1;
    for (int k0 = gl; k0 < ld; k0 += lpr) {
        const int64_t idx = (int64_t)q * ld + k0;
        const int k = min(k0, kmax);
        double g[STB_MAXM];
#pragma unroll
        for (int m = 0; m < STB_MAXM; ++m) g[m] = B.g[m] ? ld_stream(B.g[m] + idx) : 0.0;
#pragma unroll
        for (int s = 0; s < STB_MAXC; ++s) {
            if (s < ns) {
                const int m = B.slot_mult[s];
                const double gg = m == 0 ? g[0] : (m == 1 ? g[1] : g[2]);
                acc[s] = fma(gg, zU[B.slot_col[s] * nU + k], acc[s]);
            }
        }
    }
}

// Fields of the rows of slice `slice` of type T gathered from the plane vector z (global memory
// or a shared-memory copy), multiplied by `scale`.  On return every lane of a row holds the
// row's reduced sums in Fq[0..nf[T]): partner-S slots first, then partner-A slots.  One
// (non-inlined) instance serves the RHS, Jv and integrator kernels.
// Return value: lane `gl` (lane index inside the row's group of lpr lanes) holds field slot gl
// (0 for gl >= nf[T]) - the form used by the warp-cooperative epilogues.  Fq may be null.
template <int UN, bool NC>
__device__ __forceinline__ double slice_fields_impl(const DevLattice& L, int T, int slice,
                                                    const double* z, int lane, double scale,
                                                    double* Fq) {
    int out = 0;
    double fl = 0.0;
    const int gl = lane & (L.lpr - 1);
    for (int U = 0; U < 2; ++U) {
        const PairBlock& B = L.blk[T][U];
        const int ns = B.n_slots;
        if (ns == 0) continue;
        double acc[STB_MAXC];
#pragma unroll
        for (int s = 0; s < STB_MAXC; ++s) acc[s] = 0.0;
        const double* zU = z + L.off[U];
        const int nU = L.n[U];
        if (L.dense) {
            if (nU > 0) {
                const int q = min(slice * L.R + lane / L.lpr, L.n[T] - 1);
                if (L.n_mult > 1) {
                    stream_dense_multi(B, nU, L.ld[U], L.lpr, zU, q, gl, acc);
                } else {
                    switch (ns) {
                        case 1: stream_dense_t<1, UN>(B, nU, L.ld[U], L.lpr, zU, q, gl, acc); break;
                        case 2: stream_dense_t<2, UN>(B, nU, L.ld[U], L.lpr, zU, q, gl, acc); break;
                        case 3: stream_dense_t<3, UN>(B, nU, L.ld[U], L.lpr, zU, q, gl, acc); break;
                        case 4: stream_dense_t<4, UN>(B, nU, L.ld[U], L.lpr, zU, q, gl, acc); break;
                        case 5: stream_dense_t<5, UN>(B, nU, L.ld[U], L.lpr, zU, q, gl, acc); break;
                        default: stream_dense_t<6, UN>(B, nU, L.ld[U], L.lpr, zU, q, gl, acc); break;
                    }
                }
            }
        } else {
        const int64_t l0 = B.line_ptr[slice], l1 = B.line_ptr[slice + 1];
        if (L.n_mult > 1) {
            stream_block_multi(B, nU, zU, l0, l1, lane, acc);
        } else {
            switch (ns) {
                case 1: stream_block_t<1, UN, NC>(B, nU, zU, l0, l1, lane, acc); break;
                case 2: stream_block_t<2, UN, NC>(B, nU, zU, l0, l1, lane, acc); break;
                case 3: stream_block_t<3, UN, NC>(B, nU, zU, l0, l1, lane, acc); break;
                case 4: stream_block_t<4, UN, NC>(B, nU, zU, l0, l1, lane, acc); break;
                case 5: stream_block_t<5, UN, NC>(B, nU, zU, l0, l1, lane, acc); break;
                default: stream_block_t<6, UN, NC>(B, nU, zU, l0, l1, lane, acc); break;
            }
        }
        }
        for (int o = L.lpr >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int s = 0; s < STB_MAXC; ++s)
                if (s < ns) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
        }
#pragma unroll
        for (int s = 0; s < STB_MAXC; ++s) {
            if (s < ns) {
                const double v = acc[s] * scale;
                if (Fq) Fq[out + s] = v;
                if (gl == out + s) fl = v;
            }
        }
        out += ns;
    }
    if (Fq)      // slots past the pair blocks (cooperative fields) are filled elsewhere: zero here
        for (int k = out; k < L.nf[T]; ++k) Fq[k] = 0.0;
    return fl;
}

// One non-inlined instance for the persistent integrator (instruction-cache footprint).
__device__ __noinline__ double slice_fields(const DevLattice& L, int T, int slice, const double* z,
                                            int lane, double scale, double* Fq) {
    return slice_fields_impl<STB_UN_INT, true>(L, T, slice, z, lane, scale, Fq);
}

// Same, for pair data resident in shared memory (generic loads).
__device__ __noinline__ double slice_fields_resident(const DevLattice& L, int T, int slice,
                                                     const double* z, int lane, double scale,
                                                     double* Fq) {
    return slice_fields_impl<STB_UN_INT, false>(L, T, slice, z, lane, scale, Fq);
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
    if (threadIdx.x == 0 && L.n_coop) P.C[L.n_proc] = 1.0;     // cooperative terms: w carries the strength
}

// the same with the gathered vectors read through L2 (grid integrator: other CTAs of the group wrote
// them; its group barrier is release-only, so L1 may hold stale lines)
__device__ __forceinline__ double coop_list_field_cg(const CoopList& C, int q, const double* y1, const double* y2,
                                                     const double* v1, const double* v2, int lane) {
    const int64_t e0 = C.row_ptr[q], e1 = C.row_ptr[q + 1];
    double acc = 0.0;
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        const uint32_t a = C.p1[e], b = C.p2[e];
        const double w = C.w[e];
        if (v1) acc = fma(w, fma(__ldcg(v1 + a), __ldcg(y2 + b), __ldcg(y1 + a) * __ldcg(v2 + b)), acc);
        else acc = fma(w * __ldcg(y1 + a), __ldcg(y2 + b), acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// cooperative field of one ion (see CoopList in stb_common.cuh)
__device__ __forceinline__ double coop_list_field(const CoopList& C, int q, const double* y1, const double* y2,
                                                  const double* v1, const double* v2, int lane) {
    const int64_t e0 = C.row_ptr[q], e1 = C.row_ptr[q + 1];
    double acc = 0.0;
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        const uint32_t a = C.p1[e], b = C.p2[e];
        const double w = C.w[e];
        if (v1) acc = fma(w, fma(v1[a], y2[b], y1[a] * v2[b]), acc);
        else acc = fma(w * y1[a], y2[b], acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

