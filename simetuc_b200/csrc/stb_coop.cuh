// Warp-cooperative per-ion epilogues (used when a row is handled by W = lpr >= 16 lanes).
//
// After the pair stream every lane of the row's group holds the reduced fields; instead of one
// lead lane doing the s x s block work serially out of local memory, lane j of the group owns
// state j of the ion: values are exchanged with width-W shuffles and nothing is spilled.
//   lane j < s   : state j of ion q          lane k < nf : field slot k (value `fl`)
//   lane t       : rate term t (chunks of W when there are more terms than lanes)
#pragma once
#include "stb_fields.cuh"

#define FULL 0xffffffffu

// k_j = sum_t rate_t * ([b_t == j] - [a_t == j]), rate_t supplied per term by `rate_of(t)`.
template <class RateFn>
__device__ __forceinline__ double coop_scatter_terms(const TermTable& tt, int W, int gl, RateFn rate_of) {
    double kj = 0.0;
    const int nt = tt.n_terms;
    for (int t0 = 0; t0 < nt; t0 += W) {
        const int t = t0 + gl;
        const double rate = rate_of(t < nt ? t : nt - 1, t < nt);
        const int m = nt - t0 < W ? nt - t0 : W;
        for (int u = 0; u < m; ++u) {
            const double r = __shfl_sync(FULL, rate, u, W);
            const int tu = t0 + u;
            kj += (tt.b[tu] == gl ? r : 0.0) - (tt.a[tu] == gl ? r : 0.0);
        }
    }
    return kj;
}

// (Lin_T x)_j with lane b holding x_b
__device__ __forceinline__ double coop_lin(const double* lin, int s, int W, int gl, double xj) {
    double acc = 0.0;
    const int row = gl < s ? gl : 0;
    for (int b = 0; b < s; ++b) acc = fma(lin[row * s + b], __shfl_sync(FULL, xj, b, W), acc);
    return acc;
}

// f_j = (Lin y)_j + sum_t C_t y[a_t] F[slot_t] ([b_t==j] - [a_t==j])        (odesolver.py:33-43)
__device__ __forceinline__ double coop_rhs(const DevLattice& L, const SysParams& P, int T, int W,
                                           int gl, double yj, double fl) {
    const TermTable& tt = L.terms[T];
    double out = coop_lin(P.lin[T], L.s[T], W, gl, yj);
    if (tt.n_terms) {
        out += coop_scatter_terms(tt, W, gl, [&](int t, bool ok) {
            const double ya = __shfl_sync(FULL, yj, tt.a[t], W);
            const double ft = __shfl_sync(FULL, fl, tt.fslot[t], W);
            return ok ? P.C[tt.proc[t]] * ya * ft : 0.0;
        });
    }
    return out;
}

// (J v)_j = (Lin v)_j + sum_t C_t (v[a] Fy[slot] + y[a] Fv[slot]) ([b==j] - [a==j])  (odesolver.py:46-62)
__device__ __forceinline__ double coop_jv(const DevLattice& L, const SysParams& P, int T, int W,
                                          int gl, double yj, double vj, double fyl, double fvl) {
    const TermTable& tt = L.terms[T];
    double out = coop_lin(P.lin[T], L.s[T], W, gl, vj);
    if (tt.n_terms) {
        out += coop_scatter_terms(tt, W, gl, [&](int t, bool ok) {
            const int a = tt.a[t], f = tt.fslot[t];
            const double ya = __shfl_sync(FULL, yj, a, W), va = __shfl_sync(FULL, vj, a, W);
            const double fy = __shfl_sync(FULL, fyl, f, W), fv = __shfl_sync(FULL, fvl, f, W);
            return ok ? P.C[tt.proc[t]] * fma(va, fy, ya * fv) : 0.0;
        });
    }
    return out;
}

// Coupling part only: k_j = sum_t C_t y[a_t] Fv[slot_t] ([b==j] - [a==j]); with
// Minv = (I - c Jblock)^-1 the preconditioned operator is  M^-1 (I - cJ) v = v - c Minv k.
__device__ __forceinline__ double coop_coupling(const DevLattice& L, const SysParams& P, int T, int W,
                                                int gl, double yj, double fvl) {
    const TermTable& tt = L.terms[T];
    if (!tt.n_terms) return 0.0;
    return coop_scatter_terms(tt, W, gl, [&](int t, bool ok) {
        const double ya = __shfl_sync(FULL, yj, tt.a[t], W);
        const double fv = __shfl_sync(FULL, fvl, tt.fslot[t], W);
        return ok ? P.C[tt.proc[t]] * ya * fv : 0.0;
    });
}

// Row j of M = I - c (Lin_T + sum_t C_t Fy[slot_t] (e_b - e_a) e_a^T) in lane j (a[k], k < s).
__device__ __forceinline__ void coop_block_row(const DevLattice& L, const SysParams& P, int T,
                                               int W, int gl, double fyl, double c,
                                               double (&a)[STB_MAXS]) {
    const int s = L.s[T];
    const double* lin = P.lin[T];
    const TermTable& tt = L.terms[T];
    const int row = gl < s ? gl : 0;
#pragma unroll
    for (int k = 0; k < STB_MAXS; ++k) {
        a[k] = k < s ? -c * lin[row * s + k] : 0.0;
        if (k == gl) a[k] += 1.0;
    }
    for (int t = 0; t < tt.n_terms; ++t) {
        const double kk = c * P.C[tt.proc[t]] * __shfl_sync(FULL, fyl, tt.fslot[t], W);
        const int ta = tt.a[t], tb = tt.b[t];
#pragma unroll
        for (int k = 0; k < STB_MAXS; ++k) {
            if (k == ta) a[k] += (gl == ta ? kk : 0.0) - (gl == tb ? kk : 0.0);
        }
    }
}

// Row j of Minv = M^-1 in lane j.
// Gauss-Jordan without pivoting: the matrix is a column-diagonally-dominant M-matrix (columns of
// every rate operator sum to zero, so columns of I - cJ sum to one with non-positive
// off-diagonals).  Rows live in registers (static indices), pivot rows travel by shuffle.
__device__ __forceinline__ void coop_block_inverse(const DevLattice& L, const SysParams& P, int T,
                                                   int W, int gl, double fyl, double c,
                                                   double (&ai)[STB_MAXS]) {
    const int s = L.s[T];
    double a[STB_MAXS];
    coop_block_row(L, P, T, W, gl, fyl, c, a);
#pragma unroll
    for (int k = 0; k < STB_MAXS; ++k) ai[k] = k == gl ? 1.0 : 0.0;
    for (int col = 0; col < s; ++col) {
        double acol = 0.0;
#pragma unroll
        for (int k = 0; k < STB_MAXS; ++k) acol = k == col ? a[k] : acol;
        const double piv = __shfl_sync(FULL, acol, col, W);
        const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
        const bool mine = gl == col;
        const double m = mine ? 0.0 : acol;
#pragma unroll
        for (int k = 0; k < STB_MAXS; ++k) {
            if (k < s) {
                const double pa = __shfl_sync(FULL, a[k], col, W) * inv;
                const double pi = __shfl_sync(FULL, ai[k], col, W) * inv;
                a[k] = mine ? pa : fma(-m, pa, a[k]);
                ai[k] = mine ? pi : fma(-m, pi, ai[k]);
            }
        }
    }
}

// x_j = sum_b ai[b] * r_b with lane b holding r_b (row j of the inverse in registers)
__device__ __forceinline__ double coop_apply_rows(const double (&ai)[STB_MAXS], int s, int W, double rj) {
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < STB_MAXS; ++b)
        if (b < s) acc = fma(ai[b], __shfl_sync(FULL, rj, b, W), acc);
    return acc;
}
