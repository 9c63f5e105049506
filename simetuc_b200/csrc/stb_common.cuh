// simetuc_b200 - shared device/host structures of the rate-equation path (sm_100a).
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   * state vectors are kept in PLANE layout inside the library: for ion type T (0 = S, 1 = A)
//     and state s, the populations of all ions of that type are contiguous:
//         z[off[T] + s * n[T] + i],  i = type-local ion index (lattice order within the type)
//     so that per-ion work (lane = ion) and partner gathers (lane = partner) are coalesced.
//     The reference layout (ion-major, lattice.py:279-292) only exists at the C ABI.
//   * pair lists are ion-owned ("pull"): every ion row holds all its partners of type U as
//     (partner index u32, d^-mult f64) in a sliced-ELL block: a slice is R = 32/lpr consecutive
//     rows of one type handled by one warp, `lpr` lanes cooperate on a row, and the slice is
//     stored as lines of 32 entries so that lane L reads entry [line*32 + L]: fully coalesced.
//   * all-to-all lattices (no cutoff: the reference's semantics, lattice.py:335) need no partner
//     index at all: the partners of a row are ALL ions of type U, so a block is the DENSE matrix
//     g[q * ld[U] + k] = d(q,k)^-mult (0 on the diagonal, for coincident ions and in the padding
//     of a row to ld[U] = round_up(n[U], 32)): 8 bytes per ordered pair instead of 12, and the
//     gather of the partner populations becomes a contiguous read.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/simetuc_b200.h"

#define STB_MAXC 6    // field slots (partner state, multipolarity) per (row type, partner type)
#define STB_MAXM 3    // distinct multipolarities
#define STB_MAXT 64   // rate terms per row type (a process contributes a donor and an acceptor term)
#define STB_MAXS STB_MAX_STATES
#define STB_KRYLOV 30 // GMRES basis size
#define STB_MAXORD 5
#define STB_NT 512    // threads per CTA of the integrator / rhs kernels
#define STB_NW (STB_NT / 32)

// Dynamic shared memory: EVERY `extern __shared__` declaration of the library goes through this
// macro.  nvcc places each such declaration at the end of the static variables rounded up to the
// declaration's OWN alignment, so two declarations with different alignments inside one kernel
// (say `dyn` in the kernel, `dynb` in an inlined device function, a third array elsewhere in the
// translation unit with `__align__(128)`) stop aliasing: measured, the cluster integrator then
// addressed its arrays from two bases 32 B apart ("Dynamic shared memory base").
// 16 B is what LDS.128 and `cp.async.bulk` destinations need.
#define STB_DYN_SMEM(name) extern __shared__ __align__(16) double name[]

struct PairBlock {
    const uint32_t* col;         // [lines*32] partner index, type-local (null: dense block)
    const double* g[STB_MAXM];   // [lines*32] d^-mult, or dense [n[T]][ld[U]]
    const int64_t* line_ptr;     // [n_slices(T)+1] (sliced-ELL only)
    int n_slots;
    int slot_col[STB_MAXC];      // partner state whose population is accumulated
    int slot_mult[STB_MAXC];     // which g[]
};

// rate = C[proc] * y_q[a] * F[slot];  dy_q[a] -= rate;  dy_q[b] += rate
struct TermTable {
    int n_terms;
    short fslot[STB_MAXT];       // index into the row's field array (U-major)
    short a[STB_MAXT], b[STB_MAXT], proc[STB_MAXT];
};

// Cooperative (three-ion, S-S-A) energy transfer (precalculate.py:427-599), ion-owned like the pair
// lists: every ion holds, per role it plays in a triple, the list of its two partners and the
// rate `strength * (1/(d_li d_lk)^m + 1/(d_li d_ik)^m + 1/(d_ik d_lk)^m)`.  The sum over the list
// of w * z[p1] * z[p2] is one more FIELD of the ion (slot `fslot` of its type), so that the rate
// terms, the block-Jacobi preconditioner and the Jacobian-vector product treat it like a pair
// process: rate = y_q[a] * F[fslot], moved from state a to state b.
struct CoopList {
    const int64_t* row_ptr;      // [n[T] + 1]
    const uint32_t* p1;          // type-local index of the first partner (type T1, state col1)
    const uint32_t* p2;          // ... second partner (type T2, state col2)
    const double* w;
    int T, T1, col1, T2, col2, fslot, a, b;
};

struct DevLattice {
    int n_ions, N, NF;
    int n[2], s[2], off[2];      // ions / states / plane offset per type
    int lpr, R;                  // lanes per row, rows per slice (R*lpr == 32)
    int dense;                   // 1: dense all-to-all blocks (no partner index), 0: sliced-ELL
    int ld[2];                   // dense: row pitch of a block over partners of type U
    int n_slices[2];
    int n_mult, mults[STB_MAXM];
    int n_proc;
    int nf[2], foff[2];          // field slots per row type, offset of the type's field planes
    PairBlock blk[2][2];         // [row type][partner type]
    TermTable terms[2];
    const int* plane_of_ref;     // [N]
    const int* ref_of_plane;     // [N]
    int blk_stride;              // doubles per system in the `blocks` array
    int n_coop;                  // 0, or 3 roles: S as first donor, S as second donor, A as acceptor
    CoopList coop[3];
    int quad;                    // 1: the sensitizers are stored in quads of strongly coupled ions
                                 // (type-local indices 4m .. 4m+3), see group_sensitizers()
    int grp16;                   // 1: ... and the quads in groups of four neighbouring quads (16m .. 16m+15)
    int pcg;                     // ions per block of the cluster integrator's group preconditioner (4 or 16)
};

// Rows of type T owned by CTA `rank` of a group of G CTAs.  Quad lattices split the sensitizers in
// whole preconditioner groups (L.pcg ions), so that a block never straddles two CTAs.
__host__ __device__ inline void stb_own_rows(const DevLattice& L, int T, int rank, int G, int& lo, int& cnt) {
    if (T == 0 && L.quad) {
        const int pg = L.pcg;
        const int nq = (L.n[0] + pg - 1) / pg;
        const int q_lo = (int)(((long long)nq * rank) / G), q_hi = (int)(((long long)nq * (rank + 1)) / G);
        lo = pg * q_lo;
        const int hi = pg * q_hi < L.n[0] ? pg * q_hi : L.n[0];
        cnt = hi > lo ? hi - lo : 0;
    } else {
        lo = (int)(((long long)L.n[T] * rank) / G);
        cnt = (int)(((long long)L.n[T] * (rank + 1)) / G) - lo;
    }
}
// the largest such count over the ranks
__host__ __device__ inline int stb_own_rows_max(const DevLattice& L, int T, int G) {
    if (T == 0 && L.quad) return L.pcg * (((L.n[0] + L.pcg - 1) / L.pcg + G - 1) / G);
    return (L.n[T] + G - 1) / G;
}

// per-system parameters as the kernels see them
struct SysParams {
    double lin[2][STB_MAXS * STB_MAXS];  // decay (+ absorption when pulse), row-major
    double C[STB_MAX_PROC];              // effective strengths (0 when |strength| <= 1e-8)
};

__host__ __device__ inline bool stb_active_strength(double s) {
    // the reference skips a process when np.isclose(strength, 0.0) (precalculate.py:354)
    return (s > 1e-8) || (s < -1e-8);
}
