// ET coupling build on the device: distances -> ion-owned sliced-ELL pair lists, and the
// reference-ordered interaction view (N_indices, ET columns, jac_indices).
//
// Replaces lattice.create_interaction_matrices (lattice.py:301-374: for every ion the ascending
// list of all other ions of each type with non-zero distance) and the pair part of
// precalculate._create_ET_matrices (precalculate.py:235-415: w = strength * d^-mult).
#pragma once
#include "stb_common.cuh"

struct Geo {
    const double* dist;   // [n*n] or null
    const double* pos;    // [n*3] or null
    double cell[9], inv[9];
    int periodic;
    int n;
    double d_max;         // <= 0: no cutoff
};

// minimum-image distance, same arithmetic as simetuc_b200/lattice.py:mic_distances
__host__ __device__ __forceinline__ double geo_dist(const Geo& G, int a, int b) {
    if (G.dist) return G.dist[(int64_t)a * G.n + b];
    double d0 = G.pos[3 * b] - G.pos[3 * a];
    double d1 = G.pos[3 * b + 1] - G.pos[3 * a + 1];
    double d2 = G.pos[3 * b + 2] - G.pos[3 * a + 2];
    if (!G.periodic) return sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    double f0 = d0 * G.inv[0] + d1 * G.inv[3] + d2 * G.inv[6];
    double f1 = d0 * G.inv[1] + d1 * G.inv[4] + d2 * G.inv[7];
    double f2 = d0 * G.inv[2] + d1 * G.inv[5] + d2 * G.inv[8];
    f0 -= rint(f0); f1 -= rint(f1); f2 -= rint(f2);
    d0 = f0 * G.cell[0] + f1 * G.cell[3] + f2 * G.cell[6];
    d1 = f0 * G.cell[1] + f1 * G.cell[4] + f2 * G.cell[7];
    d2 = f0 * G.cell[2] + f1 * G.cell[5] + f2 * G.cell[8];
    double best = 1e300;
    for (int i = -1; i <= 1; ++i)
        for (int j = -1; j <= 1; ++j)
            for (int k = -1; k <= 1; ++k) {
                double e0 = d0 + i * G.cell[0] + j * G.cell[3] + k * G.cell[6];
                double e1 = d1 + i * G.cell[1] + j * G.cell[4] + k * G.cell[7];
                double e2 = d2 + i * G.cell[2] + j * G.cell[5] + k * G.cell[8];
                double r2 = e0 * e0 + e1 * e1 + e2 * e2;
                best = r2 < best ? r2 : best;
            }
    return sqrt(best);
}

__device__ __forceinline__ bool geo_pair(const Geo& G, double d) {
    return d != 0.0 && (G.d_max <= 0.0 || d <= G.d_max);
}

// row lengths: len[U][ion] = number of partners of type U of ion `ion` (global ion index)
__global__ void k_count_pairs(Geo G, const uint8_t* type, int* len /*[2][n]*/) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int q = warp; q < G.n; q += nwarps) {
        int c0 = 0, c1 = 0;
        for (int k0 = 0; k0 < G.n; k0 += 32) {
            int k = k0 + lane;
            bool ok = false;
            int U = 0;
            if (k < G.n) {
                ok = geo_pair(G, geo_dist(G, q, k));
                U = type[k] != 0;
            }
            c0 += __popc(__ballot_sync(0xffffffffu, ok && U == 0));
            c1 += __popc(__ballot_sync(0xffffffffu, ok && U == 1));
        }
        if (lane == 0) { len[q] = c0; len[G.n + q] = c1; }
    }
}

struct FillArgs {
    Geo G;
    const uint8_t* type;
    const int* local_idx;       // [n] type-local index of each ion
    int lpr, R, n_mult;
    int mults[STB_MAXM];
    uint32_t* col[2][2];        // [T][U]
    double* g[2][2][STB_MAXM];
    const int64_t* line_ptr[2][2];
};

// one warp per ion row; partners in ascending lattice order (lattice.py:335-371)
__global__ void k_fill_pairs(FillArgs A) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    const Geo& G = A.G;
    for (int q = warp; q < G.n; q += nwarps) {
        int T = A.type[q] != 0;
        int ql = A.local_idx[q];
        int slice = ql / A.R, rl = ql % A.R;
        int e[2] = {0, 0};
        for (int k0 = 0; k0 < G.n; k0 += 32) {
            int k = k0 + lane;
            bool ok = false;
            int U = 0;
            double d = 0.0;
            if (k < G.n) {
                d = geo_dist(G, q, k);
                ok = geo_pair(G, d);
                U = A.type[k] != 0;
            }
#pragma unroll
            for (int UU = 0; UU < 2; ++UU) {
                unsigned m = __ballot_sync(0xffffffffu, ok && U == UU);
                if (ok && U == UU && A.col[T][UU] != nullptr) {
                    int pos = e[UU] + __popc(m & ((1u << lane) - 1u));
                    int64_t line = A.line_ptr[T][UU][slice] + pos / A.lpr;
                    int64_t idx = line * 32 + rl * A.lpr + pos % A.lpr;
                    A.col[T][UU][idx] = (uint32_t)A.local_idx[k];
                    for (int mm = 0; mm < A.n_mult; ++mm)
                        A.g[T][UU][mm][idx] = pow(d, -(double)A.mults[mm]);
                }
                e[UU] += __popc(m);
            }
        }
    }
}

// ---- cell lists for cutoff lattices (d_max > 0, positions given) -----------------------------------
// The supercell (or the bounding box of an open lattice) is cut into nc[0] x nc[1] x nc[2] cells
// no thinner than d_max (in fractional coordinates), so every partner within the cutoff sits in one
// of the 27 cells around an ion's own cell: the build is O(n * partners) instead of the O(n^2) scan
// of k_count_pairs / k_fill_pairs.  The distances come from the same geo_dist (bit-identical rates)
// and the partners of a row are written in ascending lattice order (lattice.py:335-371) like the
// scan does: candidates are collected in shared memory and placed by rank.
struct CellGrid {
    const int* cell_start;   // [n_cells + 1]
    const int* cell_ions;    // [n] ions of cell c at cell_start[c] .. cell_start[c + 1], ascending
    const int* ion_cell;     // [n] cell of every ion
    int nc[3];
    int periodic;
};

// calls f(k, valid) for the candidate partners of ion q, 32 at a time (uniform trip count per warp)
template <class F>
__device__ __forceinline__ void cells_for_each(const CellGrid& C, int q, int lane, F&& f) {
    const int c = C.ion_cell[q];
    const int cz = c % C.nc[2], cy = (c / C.nc[2]) % C.nc[1], cx = c / (C.nc[2] * C.nc[1]);
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz) {
                int x = cx + dx, y = cy + dy, z = cz + dz;
                if (C.periodic) {
                    x = (x + C.nc[0]) % C.nc[0];
                    y = (y + C.nc[1]) % C.nc[1];
                    z = (z + C.nc[2]) % C.nc[2];
                } else if (x < 0 || y < 0 || z < 0 || x >= C.nc[0] || y >= C.nc[1] || z >= C.nc[2]) {
                    continue;
                }
                const int cc = (x * C.nc[1] + y) * C.nc[2] + z;
                const int s0 = C.cell_start[cc], s1 = C.cell_start[cc + 1];
                for (int s = s0; s < s1; s += 32) {
                    const bool valid = s + lane < s1;
                    f(valid ? C.cell_ions[s + lane] : 0, valid);
                }
            }
}

__global__ void k_count_pairs_cells(Geo G, CellGrid C, const uint8_t* type, int* len /*[2][n]*/) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int q = warp; q < G.n; q += nwarps) {
        int c0 = 0, c1 = 0;
        cells_for_each(C, q, lane, [&](int k, bool valid) {
            const bool ok = valid && geo_pair(G, geo_dist(G, q, k));
            const int U = valid ? (type[k] != 0) : 0;
            c0 += __popc(__ballot_sync(0xffffffffu, ok && U == 0));
            c1 += __popc(__ballot_sync(0xffffffffu, ok && U == 1));
        });
        if (lane == 0) { len[q] = c0; len[G.n + q] = c1; }
    }
}

#define STB_CELL_CAP 1024     // accepted partners of one row kept in shared memory (per warp)

__global__ void __launch_bounds__(128) k_fill_pairs_cells(FillArgs A, CellGrid C) {
    __shared__ unsigned ks[4][STB_CELL_CAP];      // ion index | partner type << 31
    __shared__ double ds[4][STB_CELL_CAP];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    const Geo& G = A.G;
    for (int q = warp; q < G.n; q += nwarps) {
        const int T = A.type[q] != 0;
        const int ql = A.local_idx[q];
        const int slice = ql / A.R, rl = ql % A.R;
        int cnt = 0;
        bool overflow = false;
        cells_for_each(C, q, lane, [&](int k, bool valid) {
            double d = 0.0;
            bool ok = false;
            if (valid) {
                d = geo_dist(G, q, k);
                ok = geo_pair(G, d);
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            const int pos = cnt + __popc(m & ((1u << lane) - 1u));
            if (ok && pos < STB_CELL_CAP) {
                ks[wib][pos] = (unsigned)k | (A.type[k] != 0 ? 0x80000000u : 0u);
                ds[wib][pos] = d;
            }
            cnt += __popc(m);
            if (cnt > STB_CELL_CAP) overflow = true;
        });
        __syncwarp();
        if (overflow) {
            // a row with more partners than the buffer holds: the ordered scan of k_fill_pairs
            int e[2] = {0, 0};
            for (int k0 = 0; k0 < G.n; k0 += 32) {
                const int k = k0 + lane;
                bool ok = false;
                int U = 0;
                double d = 0.0;
                if (k < G.n) {
                    d = geo_dist(G, q, k);
                    ok = geo_pair(G, d);
                    U = A.type[k] != 0;
                }
#pragma unroll
                for (int UU = 0; UU < 2; ++UU) {
                    const unsigned m = __ballot_sync(0xffffffffu, ok && U == UU);
                    if (ok && U == UU && A.col[T][UU] != nullptr) {
                        const int pos = e[UU] + __popc(m & ((1u << lane) - 1u));
                        const int64_t line = A.line_ptr[T][UU][slice] + pos / A.lpr;
                        const int64_t idx = line * 32 + rl * A.lpr + pos % A.lpr;
                        A.col[T][UU][idx] = (uint32_t)A.local_idx[k];
                        for (int mm = 0; mm < A.n_mult; ++mm) A.g[T][UU][mm][idx] = pow(d, -(double)A.mults[mm]);
                    }
                    e[UU] += __popc(m);
                }
            }
            continue;
        }
        // place every accepted partner by its rank among the partners of its type (ascending ion index)
        for (int e = lane; e < cnt; e += 32) {
            const unsigned me = ks[wib][e];
            const int UU = me >> 31;
            if (A.col[T][UU] == nullptr) continue;
            int pos = 0;
            for (int o = 0; o < cnt; ++o) {
                const unsigned other = ks[wib][o];
                pos += ((other >> 31) == (unsigned)UU) && other < me;
            }
            const int k = (int)(me & 0x7fffffffu);
            const double d = ds[wib][e];
            const int64_t line = A.line_ptr[T][UU][slice] + pos / A.lpr;
            const int64_t idx = line * 32 + rl * A.lpr + pos % A.lpr;
            A.col[T][UU][idx] = (uint32_t)A.local_idx[k];
            for (int mm = 0; mm < A.n_mult; ++mm) A.g[T][UU][mm][idx] = pow(d, -(double)A.mults[mm]);
        }
        __syncwarp();
    }
}

struct FillDenseArgs {
    Geo G;
    const uint8_t* type;
    const int* local_idx;
    int n_mult;
    int mults[STB_MAXM];
    int ld[2];
    double* g[2][2][STB_MAXM];  // [T][U][m] -> [n_T][ld[U]], zero-initialised; null = unused block
};

// dense all-to-all blocks: g[T][U][ql * ld[U] + kl] = d(q,k)^-mult, one warp per ion row
__global__ void k_fill_dense(FillDenseArgs A) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    const Geo& G = A.G;
    for (int q = warp; q < G.n; q += nwarps) {
        const int T = A.type[q] != 0;
        const int64_t ql = A.local_idx[q];
        for (int k = lane; k < G.n; k += 32) {
            const int U = A.type[k] != 0;
            if (A.g[T][U][0] == nullptr) continue;
            const double d = geo_dist(G, q, k);
            if (!geo_pair(G, d)) continue;
            const int64_t idx = ql * A.ld[U] + A.local_idx[k];
            for (int mm = 0; mm < A.n_mult; ++mm)
                A.g[T][U][mm][idx] = pow(d, -(double)A.mults[mm]);
        }
    }
}

struct ExportArgs {
    const DevLattice* L;
    const uint8_t* type;
    const int* local_idx;
    const int* gs_ref;          // [n] ground-state index of each ion, reference layout
    const int* gs_of_local[2];  // [n_T] the same, by type-local index
    const int* local_of_rank[2]; // [n_T] type-local index of the r-th ion of the type in lattice
                                //       order (null: identity) - dense blocks only
    const int* len;             // [2][n]
    const int64_t* offset;      // [n] first interaction of each ion
    const stb_process* procs;   // device copy, strengths already replaced by the system's
    int n_proc, n;
    uint32_t* N_indices;        // [M,2]
    int32_t* et_rows;           // [M,4]
    double* et_w;               // [M]
    uint32_t* jac;              // [2M,3]
};

// The reference's interaction order (precalculate.py:349-392): ions in lattice order; for an
// ion of type T first the (T,T) processes in dict order, each over all same-type partners,
// then the (T, other) processes over all other-type partners.
__global__ void k_export(ExportArgs A) {
    const DevLattice& L = *A.L;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int q = warp; q < A.n; q += nwarps) {
        int T = A.type[q] != 0;
        int ql = A.local_idx[q];
        int slice = ql / L.R, rl = ql % L.R;
        int64_t r = A.offset[q];
        int gs = A.gs_ref[q];
        for (int ord = 0; ord < 2; ++ord) {
            int U = ord == 0 ? T : 1 - T;
            int cnt = A.len[U * A.n + q];
            const PairBlock& B = L.blk[T][U];
            for (int p = 0; p < A.n_proc; ++p) {
                const stb_process pr = A.procs[p];
                if (pr.donor_type != T || pr.partner_type != U || !stb_active_strength(pr.strength))
                    continue;
                int mi = 0;
                for (int m = 0; m < L.n_mult; ++m) if (L.mults[m] == pr.mult) mi = m;
                // dense blocks: the partner list is every column with a non-zero rate, compacted
                // in ascending order (np.nonzero(dist), lattice.py:335); sliced-ELL: stored lists
                const int span = L.dense ? ((L.n[U] + 31) & ~31) : ((cnt + 31) & ~31);
                int seen = 0;
                for (int e0 = 0; e0 < span; e0 += 32) {
                    int e = e0 + lane;
                    int64_t idx = 0;
                    int kl = 0;
                    bool ok;
                    if (L.dense) {
                        // partners in LATTICE order (the r-th ion of type U), wherever it is stored
                        kl = e < L.n[U] ? (A.local_of_rank[U] ? A.local_of_rank[U][e] : e) : 0;
                        idx = (int64_t)ql * L.ld[U] + kl;
                        ok = e < L.n[U] && B.g[0][idx] != 0.0;
                        const unsigned m = __ballot_sync(0xffffffffu, ok);
                        e = seen + __popc(m & ((1u << lane) - 1u));
                        seen += __popc(m);
                    } else {
                        ok = e < cnt;
                        if (ok) {
                            int64_t line = B.line_ptr[slice] + e / L.lpr;
                            idx = line * 32 + rl * L.lpr + e % L.lpr;
                            kl = B.col[idx];
                        }
                    }
                    if (!ok) continue;
                    int gk = A.gs_of_local[U][kl];
                    double w = B.g[mi][idx] * pr.strength;   // dist**(-mult)*strength (:289)
                    int64_t o = r + e;
                    if (A.N_indices) {
                        A.N_indices[2 * o] = gs + pr.ii;
                        A.N_indices[2 * o + 1] = gk + pr.fi;
                    }
                    if (A.et_rows) {
                        A.et_rows[4 * o] = gs + pr.ii;
                        A.et_rows[4 * o + 1] = gs + pr.if_;
                        A.et_rows[4 * o + 2] = gk + pr.fi;
                        A.et_rows[4 * o + 3] = gk + pr.ff;
                    }
                    if (A.et_w) A.et_w[o] = w;
                    if (A.jac) {
                        uint32_t* j = A.jac + 6 * o;
                        j[0] = (uint32_t)o; j[1] = gs + pr.ii; j[2] = gk + pr.fi;
                        j[3] = (uint32_t)o; j[4] = gk + pr.fi; j[5] = gs + pr.ii;
                    }
                }
                r += cnt;
            }
        }
    }
}
