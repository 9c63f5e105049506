/*
 * simetuc_b200 - C ABI of the B200-native microscopic rate-equation path.
 *
 * This is the drop-in boundary: plain C, caller-owned pointers and sizes, no
 * torch / numpy types.  Every entry point names the reference interface it replaces
 * (paths relative to /root/reference/simetuc).  The Python host side
 * (`simetuc_b200/_lib.py`) binds these with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - ion types: 0 = sensitizer (S), 1 = activator (A)          (util.py:82-87)
 *   - state vectors at this boundary use the REFERENCE layout: ion-major, ions in lattice
 *     order, ion q at y[gs_q .. gs_q + states(type_q))           (lattice.py:259-298)
 *   - all floating point is IEEE float64; indices as documented per call
 *   - every function returns 0 on success, a negative STB_E* code on failure; the message is
 *     available from stb_last_error() (thread-local).  Solver non-convergence is NOT an error
 *     (odesolver.py:129-134): stb_integrate returns 0 and reports it in stb_stats.status.
 *   - pointers are HOST pointers unless the parameter is documented "device" and the call has
 *     the _dev suffix.
 */
#ifndef SIMETUC_B200_H
#define SIMETUC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STB_OK 0
#define STB_EINVAL (-1)   /* bad argument (maps to ValueError / lattice.LatticeError) */
#define STB_ECUDA (-2)    /* CUDA runtime failure */
#define STB_ENOMEM (-3)
#define STB_ESTATE (-4)   /* call made in the wrong order (e.g. batch not set) */

#define STB_MAX_STATES 12 /* states per ion type */
#define STB_MAX_PROC 32   /* pair ET processes */

/* One pair energy-transfer process: util.py:235-292 (EneryTransferProcess).
 * (ii, fi, if, ff) = process.indices: initial state of donor / partner, final state of
 * donor / partner.  Rate of one interaction = strength * d^(-mult) * y[donor ii] * y[partner fi]
 * (precalculate.py:277-294). */
typedef struct stb_process {
    int32_t donor_type;
    int32_t partner_type;
    int32_t ii, fi, if_, ff;
    int32_t mult;
    int32_t reserved;
    double strength;
} stb_process;

/* Per-system integration report (replaces the log messages of odesolver.py:129-134). */
typedef struct stb_stats {
    int64_t n_steps;        /* accepted BDF steps */
    int64_t n_rejected;     /* steps rejected by the error test */
    int64_t n_newton_fail;  /* Newton / Krylov convergence failures (step halved) */
    int64_t n_rhs;          /* RHS evaluations (pair-stream passes with y) */
    int64_t n_jv;           /* Jacobian-vector products (pair-stream passes with v) */
    int64_t n_lin;          /* linear solves */
    int64_t n_out;          /* output rows written (== n_t on success) */
    int32_t status;         /* 0 ok, 1 = too many steps (nsteps), 2 = step size underflow,
                               3 = non-finite values */
    int32_t last_order;
    double t_reached;
    double last_h;
} stb_stats;

typedef struct stb_system stb_system; /* opaque: one doped lattice resident on one GPU */

const char* stb_last_error(void);
int stb_device_count(void);
/* Library / kernel configuration summary (arch, launch shapes) written into buf. */
int stb_describe(char* buf, int64_t buf_len);

/* ---- ET coupling build -------------------------------------------------------------------
 * Replaces lattice.create_ground_states (lattice.py:259-298), create_interaction_matrices
 * (lattice.py:301-374) and the pair part of precalculate._create_ET_matrices
 * (precalculate.py:235-415): from the ion table and geometry it builds, on the device, the
 * ion-owned pair lists (partner index + d^-mult, sliced-ELL) that RHS / Jv / integrator stream.
 *
 * Geometry: pass EITHER dist_array (n_ions x n_ions, row-major, 0 on the diagonal; the
 * reference's shared lattice input) OR positions (n_ions x 3, Angstrom) with cell (3x3 row
 * vectors of the periodic supercell, NULL = open boundaries).  d_max <= 0 pairs every ion with
 * every other ion of non-zero distance (the reference's behaviour: lattice.py:335); d_max > 0
 * applies the cutoff the reference parses but never uses (settings.py:35).
 * Processes with strength 0 may be listed (they are skipped exactly like
 * precalculate.py:353-354 until stb_set_strengths makes them non-zero). */
stb_system* stb_system_create(int32_t device, int64_t n_ions, const uint8_t* ion_type,
                              int32_t n_S_states, int32_t n_A_states,
                              const double* dist_array, const double* positions,
                              const double* cell, double d_max,
                              int32_t n_proc, const stb_process* procs);
void stb_system_destroy(stb_system* sys);

/* Sizes: out[0]=n_ions, [1]=n_S, [2]=n_A, [3]=N (states), [4]=P (stored ordered pairs),
 * [5]=bytes of pair data streamed per RHS pass, [6]=stored pair slots (padding included),
 * [7]=lanes per row.  out[5] / out[6] = bytes per stored pair: 8 for the dense all-to-all
 * blocks (d_max <= 0), 12 for the indexed lists of a cutoff lattice (one multipolarity). */
int stb_system_info(const stb_system* sys, int64_t out[8]);

/* index_S_i / index_A_j (int64[n_ions], -1 = other type) and initial_population
 * (uint64[N]) exactly as lattice.create_ground_states returns them. */
int stb_ground_states(const stb_system* sys, int64_t* index_S_i, int64_t* index_A_j,
                      uint64_t* initial_population);

/* ---- parameter batch ----------------------------------------------------------------------
 * n_sys ODE systems share the lattice; they differ in their per-type linear blocks
 * (absorption scales with power: precalculate.py:99-120) and ET strengths (optimiser
 * candidates: optimize.py:84-85).  New systems start with zero blocks and the strengths given
 * at creation. */
int stb_set_batch(stb_system* sys, int32_t n_sys);
/* which: 0 = decay block (precalculate.py:184-213), 1 = absorption block
 * (precalculate.py:96-120, already multiplied by power_dens).  blk_S is sS x sS row-major,
 * blk_A is sA x sA (either may be NULL when that type has no states).  system = -1: all. */
int stb_set_blocks(stb_system* sys, int32_t system, int32_t which, const double* blk_S,
                   const double* blk_A);
int stb_set_strengths(stb_system* sys, int32_t system, const double* strengths /*[n_proc]*/);

/* ---- reference-layout view of the pair ET operator ------------------------------------------
 * Number M of interactions of `system` in the reference's sense (precalculate.py:296,858). */
int64_t stb_num_interactions(const stb_system* sys, int32_t system);
/* The interaction list in the reference's exact order (precalculate.py:349-392):
 *   N_indices   uint32[M,2]  (precalculate.py:293-294,413)
 *   et_rows     int32[M,4]   rows of the 4 entries of ET column r: donor ii, donor if,
 *                            partner fi, partner ff (precalculate.py:277-283)
 *   et_w        double[M]    w_r = strength * d^-mult; column r holds (-w,+w,-w,+w) (:289-291)
 *   jac_indices uint32[2M,3] (precalculate.py:603-623)
 * Any output pointer may be NULL. */
int stb_export_interactions(const stb_system* sys, int32_t system, uint32_t* N_indices,
                            int32_t* et_rows, double* et_w, uint32_t* jac_indices);

/* ---- RHS and Jacobian-vector product ----------------------------------------------------------
 * stb_rhs: odesolver._rate_eq (pulse=0, odesolver.py:33-43) / _rate_eq_pulse (pulse=1, :65-71)
 * for systems [first, first+n): y, dy are [n, N] in reference layout.
 * stb_jv:  (_jac_rate_eq[_pulse](y) @ v) without forming the dense Jacobian
 * (odesolver.py:46-62, 74-80). */
int stb_rhs(stb_system* sys, int32_t first, int32_t n, int32_t pulse, const double* y,
            double* dy);
int stb_jv(stb_system* sys, int32_t first, int32_t n, int32_t pulse, const double* y,
           const double* v, double* jv);
/* Same with DEVICE pointers in the library's internal plane layout (see stb_to_planes):
 * used by bench.py to time the kernels with operands resident in HBM.  Runs `repeat` launches
 * back to back on the library stream and returns the mean kernel time in ms (CUDA events). */
int stb_rhs_dev(stb_system* sys, int32_t first, int32_t n, int32_t pulse, const double* y_planes,
                double* dy_planes, int32_t repeat, float* ms_per_launch);
int stb_jv_dev(stb_system* sys, int32_t first, int32_t n, int32_t pulse, const double* y_planes,
               const double* v_planes, double* jv_planes, int32_t repeat, float* ms_per_launch);
/* Reference layout (host) -> internal plane layout (device buffer owned by the caller,
 * n*N doubles) and back. */
int stb_to_planes(const stb_system* sys, int32_t n, const double* y_host, double* y_planes_dev);
int stb_from_planes(const stb_system* sys, int32_t n, const double* y_planes_dev, double* y_host);

/* ---- stiff integrator ------------------------------------------------------------------------
 * Replaces odesolver._solve_ode / solve_pulse / solve_relax (odesolver.py:85-209) and scipy's
 * VODE/LSODA: variable-order (1-5) BDF, Jacobian-free Newton-Krylov (GMRES) with block-Jacobi
 * preconditioners (per ion; over quads or groups of 16 strongly coupled sensitizers), entirely on the
 * device.  Which integrator (one thread-block cluster per system / the persistent grid) and which blocks
 * are used is the library's choice per call (INTEGRATION.md lists the switches); the cluster integrator may
 * abandon a launch whose quads leave too many Krylov iterations and solve it again with the groups - the
 * caller sees one result, the statistics of the launch that produced it, and a device time that includes
 * the abandoned work.
 *   t        double[n_t]           output times, t[0] = start (shared by the batch)
 *   y0       double[n, N]          initial populations (reference layout)
 *   y_out    double[n, n_t, N]     row 0 = y0 verbatim (odesolver.py:107); may be NULL
 *   avg_out  double[n, n_t, sS+sA] per-state ion averages clipped at 0
 *                                  (simulations.py:100-128); may be NULL
 *   nsteps   max internal steps between two output times (odesolver.py:89,101)
 * Rows not reached stay zero, as in the reference (odesolver.py:97). */
int stb_integrate(stb_system* sys, int32_t first, int32_t n, int32_t pulse, int64_t n_t,
                  const double* t, const double* y0, double rtol, double atol, int64_t nsteps,
                  double* y_out, double* avg_out, stb_stats* stats /*[n]*/);
/* Same with DEVICE buffers (bench.py: operands resident in HBM): t_dev [n_t], y0 in the internal
 * plane layout [n, N] (see stb_to_planes), y_out_dev [n, n_t, N] in reference layout (may be
 * NULL), avg_out_dev [n, n_t, sS+sA] (may be NULL).  stats is a HOST pointer. */
int stb_integrate_dev(stb_system* sys, int32_t first, int32_t n, int32_t pulse, int64_t n_t,
                      const double* t_dev, const double* y0_planes_dev, double rtol, double atol,
                      int64_t nsteps, double* y_out_dev, double* avg_out_dev, stb_stats* stats);
/* Pulse followed by relaxation in ONE call: what Simulations.simulate_dynamics does with
 * solve_pulse then solve_relax (simulations.py:1125-1160), and every pulsed sweep with it
 * (simulations.py:1294-1305, 1356-1372; optimize.py:84-91).  The state at t_pulse[n_tp-1] starts the
 * relaxation on the device (no host round trip between the two solves).
 *   t_pulse  double[n_tp]  (n_tp >= 2; the reference uses linspace(0, tf_p, 2), simulations.py:1129)
 *   t_sol    double[n_t]   relaxation output times, t_sol[0] = end of the pulse (:1147)
 *   y0       double[n, N]  initial populations (reference layout)
 *   y_pulse_out double[n, n_tp, N], y_out double[n, n_t, N], avg_out double[n, n_t, sS+sA]:
 *            any may be NULL (not wanted, not copied back)
 *   nsteps_pulse / nsteps: step limits of the two solves (odesolver.py:89)
 * stats_pulse / stats_relax: [n] each.  A system whose pulse did not reach its end relaxes from the
 * zero row the pulse left (odesolver.py:97); the caller sees it in stats_pulse[k].status. */
int stb_integrate_dynamics(stb_system* sys, int32_t first, int32_t n, int64_t n_tp,
                           const double* t_pulse, int64_t n_t, const double* t_sol, const double* y0,
                           double rtol, double atol, int64_t nsteps_pulse, int64_t nsteps,
                           double* y_pulse_out, double* y_out, double* avg_out,
                           stb_stats* stats_pulse, stb_stats* stats_relax);
/* Same with DEVICE buffers (bench.py: operands resident in HBM): t_pulse_dev [n_tp], t_sol_dev [n_t],
 * y0 in the internal plane layout [n, N] (left untouched), outputs in reference layout (any may be
 * NULL).  stats_* are HOST pointers. */
int stb_integrate_dynamics_dev(stb_system* sys, int32_t first, int32_t n, int64_t n_tp,
                               const double* t_pulse_dev, int64_t n_t, const double* t_sol_dev,
                               const double* y0_planes_dev, double rtol, double atol,
                               int64_t nsteps_pulse, int64_t nsteps, double* y_pulse_out_dev,
                               double* y_out_dev, double* avg_out_dev, stb_stats* stats_pulse,
                               stb_stats* stats_relax);
/* ---- optimisation objective on the device ------------------------------------------------------
 * Replaces DynamicsSolution._get_background / _interpolate_sim_data / _calc_errors
 * (simulations.py:425-527) for batches of candidate parameter sets (optimize.py:79-107): the
 * emission curves never leave the device, only [n, sS+sA] deviations come back.
 * stb_set_experiment: the experimental decay curve of state kind `kind` (0..sS+sA-1, S states
 * first): n_e points (t_exp seconds, y_exp normalised as DynamicsSolution._load_exp_data does on
 * the host, simulations.py:354-422) and tail_mean = mean of the last 50 non-zero counts
 * (simulations.py:437-441).  n_e = 0 removes the curve (deviation 0, simulations.py:520-522).
 * stb_objective: errors[n, sS+sA] of the curves of the last stb_integrate_dynamics call (n = its
 * batch size): 1/mean(sim) * sqrt(mean((sim - exp * max(sim))^2)) with sim interpolated onto the
 * experimental times, the scaled background outside the simulated range. */
int stb_set_experiment(stb_system* sys, int32_t kind, int64_t n_e, const double* t_exp,
                       const double* y_exp, double tail_mean);
int stb_objective(stb_system* sys, int32_t n, double* errors);
/* Per-handle options.  "sm_budget": number of SMs one solve of this handle may occupy (default: all).
 * Callers that run several handles concurrently on one GPU - one lattice each, as in
 * Simulations.simulate_concentration_dependence (simulations.py:1356-1372) and sample_simulation
 * (:1412-1417) - give every handle its share, so that the launches are co-resident and the
 * integrator is chosen for throughput of the whole sweep rather than for the latency of one solve. */
/* "avg_cumulative" (0/1): stb_integrate_dynamics returns in avg_out the running trapezoid of the emission
 * curves over t_sol - t_sol[0] instead of the curves: the pulsed steady state of
 * DynamicsSolution.calculate_steady_state (simulations.py:579-591) for sweeps that do not keep the
 * trajectories (the ion average commutes with the time integral). */
int stb_set_option(stb_system* sys, const char* name, double value);
/* Read-only properties of a handle: "sm_budget"; "quad" (1: the sensitizers are stored in quads of
 * strongly coupled ions for the block preconditioner - an internal order, every result is in the
 * reference's order); "pc_group" (ions per preconditioner block of the cluster integrator: 4 or 16,
 * 1 without grouping); "cell_lists" (1: the d_max pair lists were built with cell lists instead of the
 * O(n^2) scan - the feature settings.py:35 parses and lattice.py:301-374 never applies). */
int stb_get_option(const stb_system* sys, const char* name, double* value);
/* Device time (ms, CUDA events) of the pulse and of the relaxation kernel of the last
 * stb_integrate_dynamics[_dev] call. */
int stb_last_timing_dynamics(const stb_system* sys, float* ms_pulse, float* ms_relax);
/* The CUDA stream (cudaStream_t) every kernel of this system is launched on, so that callers can
 * record their own events on it; and the number of kernels launched by the library so far. */
void* stb_stream(const stb_system* sys);
/* Cooperative (three-ion, S-S-A) energy transfer: precalculate._create_coop_ET_matrices
 * (precalculate.py:427-599) and _calculate_coop_jac_matrices (:626-699), the third term of
 * odesolver._rate_eq (:40-41) and of _jac_rate_eq (:56-60).
 * Geometry as in stb_system_create (the same lattice); idx6 = (ii, ki, li, if, kf, lf), the initial
 * states of the three transitions then the final ones (util.py:247-248); d_max_coop <= 0 or inf: no
 * cutoff (precalculate.py:477).  Triples are enumerated in the reference's order; rate of a triple
 * = strength * (1/(d_li d_lk)^mult + 1/(d_li d_ik)^mult + 1/(d_ik d_lk)^mult) (:499-506, 585).
 * A strength of 0 removes the process.  Call before stb_set_batch.  Systems with a cooperative
 * process are integrated by the cluster integrator (lattices whose state fits on 8 SMs); larger
 * ones are rejected with STB_EINVAL. */
int stb_set_coop(stb_system* S, const double* dist_array, const double* positions, const double* cell,
                 const int32_t* idx6, int32_t mult, double strength, double d_max_coop);
int64_t stb_coop_count(const stb_system* S);
/* coop_N_indices [Mc][3] (u32), rows of the 6 entries of every column (i32: -w at rows 0, 2, 4,
 * +w at rows 1, 3, 5), coop_w [Mc]; any pointer may be NULL. */
int stb_export_coop(const stb_system* S, uint32_t* coop_N_indices, int32_t* coop_rows, double* coop_w);

int64_t stb_launch_count(void);

/* Device buffers of destroyed handles are parked (per device, up to 8 GB) and reused by the next
 * handle instead of going through cudaFree / cudaMalloc: a sweep such as
 * `simulate_power_dependence` (simulations.py:1277-1316) builds one system per call.  This returns
 * the parked memory to the driver; the result is the number of bytes released. */
int64_t stb_release_cached_memory(void);
/* Device time (ms, CUDA events) of the last stb_integrate call's kernels, and the number of
 * kernel launches it made. */
int stb_last_timing(const stb_system* sys, float* ms, int64_t* launches);

/* Development aid: clock64 cycles per integrator phase of the last stb_integrate call (summed
 * over systems): [0] RHS passes, [1] Krylov operator passes, [2] Krylov dot products,
 * [3] group barriers / reductions, [4] vector updates, [5] Givens, [7] rest;
 * out[8] = CTAs per group, out[9] = groups. */
int stb_last_profile(const stb_system* sys, double out[10]);

#ifdef __cplusplus
}
#endif
#endif /* SIMETUC_B200_H */
