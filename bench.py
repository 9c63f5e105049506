#!/usr/bin/env python
"""Benchmark of the microscopic rate-equation path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): `config_file_Yb_Tm.cfg` as shipped
(beta-NaYF4, N_uc=15, 5 % Yb, 0.5 % Tm, 9 Tm states, NIR_980 pulse 10 ns) with a 32-point
power-dependence sweep, 1e6..1e13 W/cm2 log-spaced: every point is one ODE system (pulse, then
relaxation on 1000 log-spaced output times, rtol 1e-8 / atol 1e-15).  One "step" = one pass of
the hot path over that batch: the 32 systems of one lattice solved on one GPU.

Multi-GPU (one process per GPU, no data-path collective):
  * `value` / `scaling: weak`: every rank solves the 32-power sweep of the SAME lattice realisation
    (seed 2000; what an optimiser batch or a power sweep replicated over the GPUs looks like), so
    the shards are equal by construction; `per_rank_ms` shows any imbalance;
  * `strong`: ONE 32-power sweep dealt round-robin over the ranks (`system_id % N`), each rank's
    emission curves gathered on rank 0 over NCCL - the literal north_star target.

At N = 1 the line also carries `configs`: sub-records for BASELINE configs C1, C3, C4, C5
(`--no-configs` skips them).  Printed JSON (rank 0): see DESIGN.md "Measurement".
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'ODE systems solved/sec (Yb-Tm dynamics); RHS+Jv HBM GB/s vs 8 TB/s peak'
N_POWERS = 32
N_UC = 15
RTOL, ATOL = 1e-8, 1e-15
SEED = 2000


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0}, 'fallback (B200_PROFILING.md)'


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (`profiles/r02_traffic.json`, written by `tools/ncu_traffic.py` from the `.ncu-rep`)."""
    path = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


def make_cte(power=1e7):
    from simetuc_b200 import configs
    return configs.yb_tm_cte(N_uc=N_UC, S_conc=5.0, A_conc=0.5, power_dens=power, t_pulse=1e-8,
                             rtol=RTOL, atol=ATOL, N_steps=1000)


def powers_list():
    from simetuc_b200.settings import parse_power_dependence
    return parse_power_dependence([1e6, 1e13, N_POWERS])


def t_grids(cte):
    from simetuc_b200 import precalculate
    tf = float(np.round(15 * np.max(precalculate.get_lifetimes(cte)), 8))
    return (np.linspace(0, 1e-8, 2, dtype=np.float64),
            np.logspace(np.log10(1e-8), np.log10(tf), 1000, dtype=np.float64))


# ------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:   # pragma: no cover
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.h)
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:   # pragma: no cover
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


# ------------------------------------------------------------------------------------- CPU arms
def _single_thread():
    """LSODA's dense LU goes through LAPACK: pin the BLAS pool to one thread so that `cores`
    means what it says (and N worker processes do not oversubscribe the box N-fold)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:   # pragma: no cover
        pass


_CPU_OPERANDS = {}


def _cpu_operands(seed):
    """Lattice + reference-layout operands of the sweep's lattice, built once per worker process
    (setup, like the GPU arm's resident system, is outside the timed solves; only the absorption
    matrix depends on the power)."""
    if seed not in _CPU_OPERANDS:
        from simetuc_b200 import lattice
        from tests.helpers import oracle_operands
        cte = make_cte()
        lat = lattice.generate_arrays(cte, seed=seed)
        g = {'ion_type': lat['ion_type'], 'dist_array': lat['dist_array'], 'info': np.array([2, 9])}
        _CPU_OPERANDS[seed] = oracle_operands(g, cte)       # ET build + matrices (precalculate.py:713-878)
    return _CPU_OPERANDS[seed]


def _cpu_solve_one(args):
    """One system of the sweep through the oracle port (reference callbacks + scipy LSODA)."""
    seed, power = args[:2]
    import warnings
    warnings.simplefilter('ignore')
    _single_thread()
    from oracle import oracle_np as onp
    op = _cpu_operands(seed)
    if power is None:          # warm-up job: imports + operand build only
        return None
    cte = make_cte(power)
    abs_m = onp.total_absorption_matrix(2, 9, cte['excitations'], op['index_S_i'], op['index_A_j'])
    t_pulse, t_sol = t_grids(cte)
    kw = dict(rtol=RTOL, atol=ATOL, nsteps=100000)
    y_p = onp.solve_pulse(t_pulse, op['initial_population'].astype(float), abs_m, op['decay'],
                          op['ET'], op['N_indices'], op['jac_indices'], **kw)
    y = onp.solve_relax(t_sol, y_p[-1], op['decay'], op['ET'], op['N_indices'], op['jac_indices'], **kw)
    steady = onp.cumulative_steady_state(t_sol - t_sol[0], y)
    avg = onp.avg_populations(steady, op['index_S_i'], op['index_A_j'], 2, 9)
    return avg[:, -1]


def cpu_baseline_sample():
    """Single-core oracle port on a bounded sample of the sweep: powers 0 and 16 of 32."""
    powers = powers_list()
    sample = [0, 16]
    _cpu_solve_one((SEED, None))                     # operand build outside the timed solves
    t0 = time.time()
    for k in sample:
        _cpu_solve_one((SEED, powers[k]))
    dt = time.time() - t0
    return {'value': len(sample) / dt, 'unit': 'solves/s', 'cores': 1, 'kind': 'port',
            'sample': 'powers #0 and #16 of the 32-point sweep (pulse + 1000-point relaxation, '
                      'scipy LSODA rtol 1e-8 through the oracle port; lattice and ET build '
                      'outside the timed solves), %.1f s' % dt}


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port: `oracle/oracle_np.py` driving
    scipy LSODA) on ALL host cores.  One step = one pass over the sweep (all 32 powers; every
    other power on boxes with fewer than 24 cores); the systems of all steps go through one work
    queue, so no core waits for the slowest power of a step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, len(os.sched_getaffinity(0)))
    powers = powers_list()
    idx = list(range(N_POWERS)) if cores >= 24 else list(range(0, N_POWERS, 2))
    # longest first (high powers take the most steps): the tail of the queue stays short
    jobs = [(SEED, powers[i]) for _ in range(args.steps) for i in sorted(idx, reverse=True)]
    # fresh interpreters (a forked copy of a process that already runs BLAS / OpenMP threads can
    # deadlock), one BLAS thread each so that `cores` means what it says
    for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[var] = '1'
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        # warm-up: every worker imports scipy and builds the lattice + ET operands (setup is outside
        # the timed region in both arms)
        pool.map(_cpu_solve_one, [(SEED, None)] * (4 * cores), chunksize=1)
        t0 = time.time()
        for _ in pool.imap_unordered(_cpu_solve_one, jobs, chunksize=1):
            pass
        dt = time.time() - t0
    value = len(jobs) / dt
    sample = ('%d of the 32 powers per step (%s), %d worker processes fed from one queue over all '
              'steps, scipy LSODA rtol 1e-8 through the oracle port; lattice + ET build outside the '
              'timed region' % (len(idx), 'all' if len(idx) == N_POWERS else 'every other one', cores))
    line = {'metric': METRIC, 'impl': 'reference', 'value': value, 'unit': 'solves/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(1),
            'cpu_baseline': {'value': value, 'unit': 'solves/s', 'cores': cores, 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': value, 'unit': 'solves/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def workload_config(n_gpus):
    return {'workload': 'C2: config_file_Yb_Tm.cfg power dependence, 32 powers 1e6..1e13 W/cm2, '
                        'N_uc=15 5%Yb 0.5%Tm (9 Tm states), pulsed 10 ns + 1000-point relaxation, '
                        'rtol 1e-8 atol 1e-15; the same lattice realisation (seed 2000) on every GPU',
            'systems_per_step_per_gpu': N_POWERS, 'n_gpus': n_gpus,
            'l2': 'L2 flushed (256 MiB write) between timed steps (the flush itself is outside the '
                  'per-step CUDA events); the pair data of this lattice (0.5 MB) lives in shared '
                  'memory inside a solve by construction',
            'parallelism': 'independent systems, %d sweep(s) of 32, no collective' % n_gpus}


# ------------------------------------------------------------------------------------- GPU arm
def build_sweep_system(local, power_idx):
    """Device-resident system of the bench lattice with one ODE system per power in `power_idx`."""
    from simetuc_b200 import lattice, precalculate
    cte = make_cte()
    powers = powers_list()
    lat = lattice.generate(cte, seed=SEED)
    system = precalculate.build_system(cte, lat, device=local, n_sys=len(power_idx))
    for slot, k in enumerate(power_idx):
        for exc in cte['excitations']['NIR_980']:
            exc.power_dens = powers[k]
        system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=slot)
    return cte, lat, system


class DeviceSweep:
    """The hot path with every operand resident in HBM: pulse launch, hand-over of the last pulse
    row, relaxation launch, emission curves and their time integral, all on the library stream."""

    def __init__(self, system, n, dev, t_pulse, t_sol):
        import torch
        self.torch, self.system, self.n = torch, system, n
        N, kinds = system.N, system.sS + system.sA
        f64 = dict(dtype=torch.float64, device=dev)
        self.ext = torch.cuda.ExternalStream(system.stream, device=dev)
        self.d_tp = torch.tensor(t_pulse, **f64)
        self.d_ts = torch.tensor(t_sol, **f64)
        self.n_t = len(t_sol)
        _, _, pop = system.ground_states()
        self.d_y0 = torch.empty(n, N, **f64)
        system.to_planes(np.tile(pop.astype(np.float64), (n, 1)), self.d_y0.data_ptr())
        self.d_yp = torch.empty(n, 2, N, **f64)
        self.d_ys = torch.empty(n, self.n_t, N, **f64)
        self.d_avg = torch.empty(n, self.n_t, kinds, **f64)
        self.d_dt = torch.tensor(np.diff(t_sol - t_sol[0]), **f64)[None, :, None]

    def step(self):
        torch, system, n = self.torch, self.system, self.n
        st_p, st_r = system.integrate_dynamics_dev(
            self.d_tp.data_ptr(), 2, self.d_ts.data_ptr(), self.n_t, self.d_y0.data_ptr(), n, RTOL, ATOL,
            y_pulse_ptr=self.d_yp.data_ptr(), y_out_ptr=self.d_ys.data_ptr(),
            avg_out_ptr=self.d_avg.data_ptr())
        ms_p, ms_r = system.last_timing_dynamics()
        with torch.cuda.stream(self.ext):
            # pulsed steady state: area under the emission curves (simulations.py:579-591)
            steady = (0.5 * (self.d_avg[:, 1:, :] + self.d_avg[:, :-1, :]) * self.d_dt).sum(dim=1)
        return ms_p, ms_r, st_p, st_r, steady


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from simetuc_b200 import distributed, precalculate
    from simetuc_b200.simulations import Simulations

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL's banner ("NCCL version ...") goes to stdout by default: keep stdout for the JSON line
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)
    peaks, peak_src = load_peaks()
    warmup = max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- setup (untimed): the sweep's lattice, device-resident system --------------------------
    cte, lat, system = build_sweep_system(local, list(range(N_POWERS)))
    powers = powers_list()
    N, kinds = system.N, system.sS + system.sA
    t_pulse, t_sol = t_grids(cte)
    sweep = DeviceSweep(system, N_POWERS, dev, t_pulse, t_sol)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    ms_pulse, ms_relax, passes, stats_last = [], [], [], None

    for _ in range(warmup):
        sweep.step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = system.lib.stb_launch_count()
    # the K timed steps: CUDA events on the library stream around every step; the L2 flush between
    # two steps is not part of a step
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    t0 = time.time()
    for ev_a, ev_b in evs:
        with torch.cuda.stream(sweep.ext):
            flush.zero_()
        ev_a.record(sweep.ext)
        ms_p, ms_r, st_p, st_r, steady = sweep.step()
        ev_b.record(sweep.ext)
        ms_pulse.append(ms_p)
        ms_relax.append(ms_r)
        passes.append((sum(s['n_rhs'] for s in st_p + st_r), sum(s['n_jv'] for s in st_p + st_r)))
        stats_last = (st_p, st_r)
    barrier()
    wall = time.time() - t0
    dev_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    launches = system.lib.stb_launch_count() - launches0
    clocks = sampler.stop()
    ok = all(s['status'] == 0 for s in stats_last[0] + stats_last[1])

    # ---- end to end through the public API with host buffers ----------------------------------
    sim = Simulations(make_cte(), devices=[local])
    sim.cte['lattice_seed'] = SEED
    precalculate._LATTICE_CACHE[precalculate._lattice_key(sim.cte)] = lat
    for _ in range(2):
        sim.simulate_power_dependence(powers, keep_y=False)         # warm-up
    barrier()
    e2e_steps = max(1, min(args.steps, 10))
    t1 = time.time()
    for _ in range(e2e_steps):
        sol = sim.simulate_power_dependence(powers, keep_y=False)
        intens = sol.populations                                   # host result of the step
    torch.cuda.synchronize()
    e2e_wall = time.time() - t1
    # per step: per-system blocks (decay + absorption, both ion types), initial populations,
    # both time grids in; emission curves and solver statistics out
    h2d = N_POWERS * 2 * 8 * (4 + 81) + N_POWERS * N * 8 + (2 + 1000) * 8
    d2h = N_POWERS * 1000 * kinds * 8 + 2 * N_POWERS * 80
    # the device-resident and the end-to-end path must agree on the emission intensities
    conc = np.array([5.0] * system.sS + [0.5] * system.sA)
    dev_int = steady.cpu().numpy() * conc
    agree = float(np.max(np.abs(dev_int - intens) / (np.abs(intens) + 1e-300)))

    # ---- strong scaling: ONE 32-power sweep dealt over the ranks, curves gathered on rank 0 -------
    strong = None
    if world > 1:
        mine = distributed.shard_round_robin(N_POWERS, world, rank)
        _, _, sys_s = build_sweep_system(local, mine)
        sw = DeviceSweep(sys_s, len(mine), dev, t_pulse, t_sol)
        for _ in range(warmup):
            sw.step()
        barrier()
        evs_s = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                 for _ in range(args.steps)]
        ts0 = time.time()
        for ev_a, ev_b in evs_s:
            with torch.cuda.stream(sw.ext):
                flush.zero_()
            ev_a.record(sw.ext)
            _, _, _, st_s, steady_s = sw.step()
            ev_b.record(sw.ext)
        torch.cuda.synchronize()
        t_gather0 = time.time()
        gathered = distributed.gather_results(steady_s.cpu().numpy())        # final gather (NCCL)
        t_gather = time.time() - t_gather0
        barrier()
        strong_wall = time.time() - ts0
        strong_ms = float(sum(a.elapsed_time(b) for a, b in evs_s))
        if rank == 0:
            full = distributed.unshard_round_robin(gathered, N_POWERS) * conc
            strong_agree = float(np.max(np.abs(full - intens) / (np.abs(intens) + 1e-300)))
        tt = torch.tensor([strong_ms, strong_wall * 1e3, t_gather * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            s_ms, s_wall, s_gather = [float(v) for v in tt.tolist()]
            strong = {'what': 'one 32-power sweep sharded system_id %% N over %d GPUs, steady-state '
                              'intensities gathered on rank 0 (NCCL all_gather)' % world,
                      'value': N_POWERS * args.steps / (s_ms * 1e-3), 'unit': 'solves/s',
                      'ms_per_sweep': s_ms / args.steps, 'wall_ms_per_sweep': s_wall / args.steps,
                      'gather_ms': s_gather, 'systems_per_gpu': len(mine),
                      'cluster_ctas_per_system': sys_s.last_profile()['G'],
                      'max_rel_diff_vs_single_gpu_sweep': strong_agree}
        sys_s.close()

    # ---- max over ranks ----------------------------------------------------------------------------
    times = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3], dtype=torch.float64, device=dev)
    per_rank = [dev_ms / args.steps]
    if world > 1:
        all_ms = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(all_ms, torch.tensor([dev_ms / args.steps], dtype=torch.float64, device=dev))
        per_rank = [float(v.item()) for v in all_ms]
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max, e2e_ms_max = [float(v) for v in times.tolist()]
    n_rhs = float(np.mean([p[0] for p in passes]))
    n_jv = float(np.mean([p[1] for p in passes]))
    bpp = system.bytes_per_pair                # 8: dense all-to-all blocks (f64 d^-6 per ordered pair)
    b_rhs = bpp * system.n_pairs + 16 * N      # + y read, dy written once
    b_jv = bpp * system.n_pairs + 24 * N
    # dominant kernel = the relaxation launch of k_integrate_cl_t (one launch per step): its passes
    # and its own CUDA-event time
    r_rhs = float(sum(s['n_rhs'] for s in stats_last[1]))
    r_jv = float(sum(s['n_jv'] for s in stats_last[1]))
    relax_ms = float(np.mean(ms_relax))
    achieved = (r_rhs * b_rhs + r_jv * b_jv) / (relax_ms * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    traffic = load_traffic()
    value = N_POWERS * world * args.steps / (dev_ms_max * 1e-3)
    smem_peak = 128.0 * system.lib.stb_device_count() * 0 + 128.0 * 148 * (clocks['sm_mhz'] or 1965.0) * 1e6 / 1e9
    line = {
        'metric': METRIC, 'value': value, 'unit': 'solves/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': warmup, 'ms_per_step': dev_ms_max / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(world), 'clocks': clocks,
        'e2e': {'value': N_POWERS * world * e2e_steps / (e2e_ms_max * 1e-3), 'unit': 'solves/s',
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'api': 'Simulations.simulate_power_dependence(powers, keep_y=False): settings -> per-power '
                       'blocks -> pulse + relaxation (one C-ABI call, device system of the lattice '
                       'taken from the pool) -> emission curves -> steady-state intensities on the host',
                'steps': e2e_steps, 'max_rel_diff_vs_device_path': agree},
        'gpu_launches': int(launches),
        'wall_ms_per_step': wall_ms_max / args.steps,
        'per_rank_ms': per_rank,
        'solver': {'ok': bool(ok), 'rhs_passes_per_step': n_rhs, 'jv_passes_per_step': n_jv,
                   'steps_min_max': [min(s['n_steps'] for s in stats_last[1]),
                                     max(s['n_steps'] for s in stats_last[1])],
                   'pulse_ms': float(np.mean(ms_pulse)), 'relax_ms': relax_ms,
                   'us_per_bdf_step_slowest_system': 1e3 * relax_ms / max(s['n_steps'] for s in stats_last[1])},
        'roofline': {'bound': 'hbm',
                     'kernel': 'k_integrate_cl_t<false>, relaxation launch (cluster BDF / Newton-Krylov '
                               'integrator; its inner loop is the pair pass, pair rows resident in shared memory)',
                     'achieved': achieved, 'peak': peaks['hbm_gbs'],
                     'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'],
                     'traffic': traffic.get('k_integrate_cl_relax_dram_bytes'),
                     'traffic_source': traffic.get('source'),
                     'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': r_rhs * b_rhs + r_jv * b_jv,
                     'ms_per_launch': relax_ms,
                     'actual_bound': 'latency / shared memory: the %d-pair lattice (%.2f MB) is resident '
                                     'in the shared memory of its cluster, so the algorithmic bytes are '
                                     'served at %.1f %% of the %.0f GB/s aggregate shared-memory bandwidth '
                                     '(128 B/clk/SM); the HBM fraction is a unit of account, the HBM-bound '
                                     'kernels are in roofline_kernels'
                                     % (system.n_pairs, system.stream_bytes / 1e6,
                                        100.0 * achieved / smem_peak, smem_peak)},
    }
    if strong is not None:
        line['strong'] = strong
    if not args.no_large:
        line['roofline_kernels'] = large_lattice_roofline(local, peaks, peak_src, traffic)
    if world == 1 and not args.no_configs:
        line['configs'] = other_configs(local, peaks)
    if world == 1 and not args.no_cpu:
        line['cpu_baseline'] = cpu_baseline_sample()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _ref_of_plane(system):
    """Permutation plane index -> reference index, derived from the ABI (identity probe)."""
    import torch
    N = system.N
    probe = torch.empty(N, dtype=torch.float64, device='cuda')
    system.to_planes(np.arange(N, dtype=np.float64)[None, :], probe.data_ptr())
    return probe.cpu().numpy().astype(np.int64)


def large_lattice_roofline(device, peaks, peak_src, traffic):
    """RHS and Jv kernels on an HBM-bound lattice (SURVEY.md 8d "C4" shape: N_uc=40, 25 % Yb,
    0.3 % Tm, all ordered pairs -> 5.8e8 pairs, 4.7 GB of pair data, >> 126 MB of L2), timed with
    CUDA events over 20 back-to-back launches each."""
    import torch
    from simetuc_b200 import _lib, configs, lattice, precalculate
    cte = configs.yb_tm_cte(N_uc=40, S_conc=25.0, A_conc=0.3)
    lat = lattice.generate(cte, seed=4000)
    system = precalculate.build_system(cte, lat, device=device, n_sys=1)
    N = system.N
    y = torch.rand(1, N, dtype=torch.float64, device='cuda') * 1e-3
    v = torch.rand(1, N, dtype=torch.float64, device='cuda')
    out = torch.empty_like(y)
    ms = C.c_float()
    res = {}
    for name in ('rhs', 'jv'):
        for rep in (5, 20):
            if name == 'rhs':
                _lib.check(system.lib.stb_rhs_dev(system.handle, 0, 1, 1, y.data_ptr(), out.data_ptr(),
                                                  rep, C.byref(ms)))
            else:
                _lib.check(system.lib.stb_jv_dev(system.handle, 0, 1, 1, y.data_ptr(), v.data_ptr(),
                                                 out.data_ptr(), rep, C.byref(ms)))
        alg = system.bytes_per_pair * system.n_pairs + (16 if name == 'rhs' else 24) * N
        gbs = alg / (ms.value * 1e-3) / 1e9
        res[name] = {'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                     'frac': gbs / peaks['hbm_gbs'], 'ms_per_launch': ms.value,
                     'algorithmic_bytes_per_launch': alg, 'frac_of_8TBs_nominal': gbs / 8000.0,
                     'traffic': traffic.get('k_op_tma_%s_dram_bytes' % name, 4.71e9),
                     'traffic_source': traffic.get('k_op_tma_source',
                                                   'profiles/r01_ncu_rhs_jv_tma.txt (ncu --set full of k_op_tma: '
                                                   'dram read+write per launch)')}
    res['workload'] = 'N_uc=40 25%%Yb 0.3%%Tm: %d ions, %d states, %d ordered pairs, %.2f GB pair stream; ' \
                      'inputs >> L2' % (system.n_ions, N, system.n_pairs, system.stream_bytes / 1e9)
    res['peak_source'] = peak_src
    system.close()
    return res


# ------------------------------------------------------------------------------------- other configs
def _dynamics_record(cte, seed, device, peaks, label):
    """One dynamics solve (pulse + relaxation) of the lattice of `cte` through the C ABI with host
    buffers; kernel time from the library's CUDA events, algorithmic bytes from its pass counters."""
    from simetuc_b200 import lattice, precalculate
    lat = lattice.generate(cte, seed=seed)
    system = precalculate.build_system(cte, lat, device=device, n_sys=1)
    _, _, pop = system.ground_states()
    tf = float(np.round(15 * np.max(precalculate.get_lifetimes(cte)), 8))
    t_pulse = np.linspace(0, 1e-8, 2)
    t_sol = np.logspace(-8, np.log10(tf), cte['simulation_params']['N_steps'])
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        _, _, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol, pop.astype(float), rtol=RTOL,
                                                        atol=ATOL, nsteps=100000, want_y=False)
        wall = time.perf_counter() - t0
        ms, _ = system.last_timing()
        if best is None or ms < best[0]:
            best = (ms, wall)
    ms, wall = best
    n_rhs = st_p[0]['n_rhs'] + st[0]['n_rhs']
    n_jv = st_p[0]['n_jv'] + st[0]['n_jv']
    b = n_rhs * (system.bytes_per_pair * system.n_pairs + 16 * system.N) + \
        n_jv * (system.bytes_per_pair * system.n_pairs + 24 * system.N)
    gbs = b / (ms * 1e-3) / 1e9
    prof = system.last_profile()
    rec = {'workload': label, 'ions': system.n_ions, 'states': system.N, 'ordered_pairs': system.n_pairs,
           'pair_data_GB': system.stream_bytes / 1e9, 'solves_per_s': 1e3 / ms, 'kernel_ms': ms,
           'e2e_ms': 1e3 * wall, 'steps': st[0]['n_steps'], 'rhs_passes': n_rhs, 'jv_passes': n_jv,
           'ok': st_p[0]['status'] == 0 and st[0]['status'] == 0,
           'kernel': 'k_integrate (persistent grid)' if system.stream_bytes > 64e6 else
                     'k_integrate_cl_t (cluster of %d CTAs)' % int(prof['G']),
           'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                        'frac': gbs / peaks['hbm_gbs']}}
    system.close()
    return rec


def other_configs(device, peaks):
    """Sub-records for BASELINE configs C1, C3, C4, C5 (C2 is the headline)."""
    import tempfile
    from simetuc_b200 import configs, optimize, precalculate
    from simetuc_b200.simulations import Simulations
    out = {}
    # C1: config_file.cfg shape (25 % Yb, 0.3 % Tm, NIR_980), dynamics, throughput scales
    for n_uc in (20, 30, 40):
        cte = configs.yb_tm_cte(N_uc=n_uc, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8,
                                rtol=RTOL, atol=ATOL, N_steps=1000)
        out['C1_N_uc_%d' % n_uc] = _dynamics_record(
            cte, 1000 + n_uc, device, peaks,
            'C1: beta-NaYF4:Yb25%%,Tm0.3%% dynamics, N_uc=%d, 980 nm pulse 10 ns + 1000-point relaxation' % n_uc)
    # C4: large single lattice with ESA excitation (NIR_800: GSA 0->3 and ESA 2->5), N_uc = 40
    cte = configs.yb_tm_cte(N_uc=40, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8,
                            rtol=RTOL, atol=ATOL, N_steps=1000)
    for exc in cte['excitations']['NIR_980']:
        exc.active = False
    for exc in cte['excitations']['NIR_800']:
        exc.active, exc.power_dens, exc.t_pulse = True, 1e7, 1e-8
    out['C4'] = _dynamics_record(cte, 4000, device, peaks,
                                 'C4: large single lattice N_uc=40 (25%Yb 0.3%Tm, 9 Tm / 2 Yb states), '
                                 'pulse + relaxation with ESA excitation NIR_800')
    # C3: concentration grid x 8 realisations, dynamics, through the public API
    concs, n_uc = [], []
    for _seed in range(8):
        for S, nuc in ((10.0, 20), (25.0, 15), (50.0, 12)):
            for A in (0.1, 0.3, 0.5, 1.0, 2.0):
                concs.append((S, A))
                n_uc.append(nuc)
    cte = configs.yb_tm_cte(N_uc=15, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8,
                            rtol=RTOL, atol=ATOL, N_steps=1000)
    cte['lattice_seed'] = 3000
    sim = Simulations(cte, devices=[device], streams_per_device=16)
    sim.simulate_concentration_dependence(concs[:2], n_uc[:2], dynamics=True, keep_y=False)    # warm-up
    t0 = time.perf_counter()
    sol = sim.simulate_concentration_dependence(concs, n_uc, dynamics=True, keep_y=False)
    dt = time.perf_counter() - t0
    ions = [int(s.cte['ions']['total']) for s in sol.solution_list]
    out['C3'] = {'workload': 'C3: Yb {10,25,50}% x Tm {0.1,0.3,0.5,1,2}% x 8 lattice realisations, dynamics, '
                             'Simulations.simulate_concentration_dependence (16 lattices in flight)',
                 'systems': len(concs), 'seconds': dt, 'solves_per_s': len(concs) / dt,
                 'ions_min_max': [min(ions), max(ions)],
                 'ok': all(s.solver_stats['relax']['status'] == 0 for s in sol.solution_list)}
    precalculate.clear_system_pool()
    # C5: 256 optimiser candidates on the C1 parity-scale lattice, objective vs decay curves
    cte = configs.yb_tm_cte(N_uc=12, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8, rtol=RTOL,
                            atol=ATOL, N_steps=1000)
    cte['lattice_seed'] = 5000
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    try:
        os.makedirs(os.path.join(tmp, 'expData', 'bNaYF4'))
        t = np.linspace(0, 0.02, 400)
        for state, (tau_r, tau_d) in (('Tm_3H4', (2e-4, 1.5e-3)), ('Tm_1G4', (1e-4, 6e-4)),
                                      ('Tm_1D2', (5e-5, 1e-4)), ('Tm_3F4', (1e-3, 8e-3))):
            np.savetxt(os.path.join(tmp, 'expData', 'bNaYF4', 'decay_%s_exc_NIR_980_25.0Yb_0.3Tm.txt' % state),
                       np.column_stack((t, np.exp(-t / tau_d) - np.exp(-t / tau_r) + 0.01)))
        os.chdir(tmp)
        sim = Simulations(cte, devices=[device])
        base = [p.strength for p in cte['optimization']['processes']]
        rng = np.random.Generator(np.random.PCG64(5000))
        cands = [list(np.array(base) * 10.0 ** rng.uniform(-1, 1, len(base))) for _ in range(256)]
        optimize.objective_batch(sim, cands[:8])                       # warm-up
        t0 = time.perf_counter()
        errors, sols = optimize.objective_batch(sim, cands)
        dt = time.perf_counter() - t0
        out['C5'] = {'workload': 'C5: 256 candidate ET-parameter sets (ETU1, ETU2, ETU3, BackET, EM x 10^U(-1,1)) '
                                 'on the C1 parity-scale lattice (N_uc=12, 25%Yb 0.3%Tm), dynamics objective '
                                 'vs 4 decay curves, optimize.objective_batch',
                     'candidates': len(cands), 'seconds': dt, 'solves_per_s': len(cands) / dt,
                     'ions': int(sols[0].cte['ions']['total']), 'states': int(sols[0].cte['states']['energy_states']),
                     'ok': bool(all(s.solver_stats['relax']['status'] == 0 for s in sols)),
                     'finite_errors': bool(np.all(np.isfinite(errors)))}
    finally:
        os.chdir(cwd)
    precalculate.clear_system_pool()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-large', action='store_true', help='skip the HBM-bound RHS/Jv measurement')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline sample')
    ap.add_argument('--no-configs', action='store_true', help='skip the C1/C3/C4/C5 sub-records')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
