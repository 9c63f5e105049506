#!/usr/bin/env python
"""Benchmark of the microscopic rate-equation path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): `config_file_Yb_Tm.cfg` as shipped
(beta-NaYF4, N_uc=15, 5 % Yb, 0.5 % Tm, 9 Tm states, NIR_980 pulse 10 ns) with a 32-point
power-dependence sweep, 1e6..1e13 W/cm2 log-spaced: every point is one ODE system (pulse, then
relaxation on 1000 log-spaced output times, rtol 1e-8 / atol 1e-15).  One "step" = one pass of
the hot path over that batch: the 32 systems of one lattice realisation solved on one GPU.
With N GPUs every rank solves the sweep of its own lattice realisation (seed 2000 + rank):
independent systems, no data-path collective, weak scaling; `value` counts the systems of all
ranks.

Printed JSON (rank 0): see the repo's DESIGN.md "Measurement".
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


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0}, 'fallback (B200_PROFILING.md)'


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


def _cpu_solve_one(args):
    """One system of the sweep through the oracle port (reference callbacks + scipy LSODA)."""
    seed, power = args[:2]
    mp_worker = len(args) > 2 and bool(args[2])
    import warnings
    warnings.simplefilter('ignore')
    if mp_worker is False:
        _single_thread()
    from oracle import oracle_np as onp
    from simetuc_b200 import lattice
    from tests.helpers import oracle_operands
    cte = make_cte(power)
    lat = lattice.generate_arrays(cte, seed=seed)
    g = {'ion_type': lat['ion_type'], 'dist_array': lat['dist_array'], 'info': np.array([2, 9])}
    op = oracle_operands(g, cte)          # ET build + matrices (precalculate.py:713-878)
    t_pulse, t_sol = t_grids(cte)
    kw = dict(rtol=RTOL, atol=ATOL, nsteps=100000)
    y_p = onp.solve_pulse(t_pulse, op['initial_population'].astype(float), op['abs'], op['decay'],
                          op['ET'], op['N_indices'], op['jac_indices'], **kw)
    y = onp.solve_relax(t_sol, y_p[-1], op['decay'], op['ET'], op['N_indices'], op['jac_indices'], **kw)
    steady = onp.cumulative_steady_state(t_sol - t_sol[0], y)
    avg = onp.avg_populations(steady, op['index_S_i'], op['index_A_j'], 2, 9)
    return avg[:, -1]


def cpu_baseline_sample():
    """Single-core oracle port on a bounded sample of the sweep: powers 0 and 16 of 32."""
    powers = powers_list()
    sample = [0, 16]
    t0 = time.time()
    for k in sample:
        _cpu_solve_one((2000, powers[k]))
    dt = time.time() - t0
    return {'value': len(sample) / dt, 'unit': 'solves/s', 'cores': 1, 'kind': 'port',
            'sample': 'powers #0 and #16 of the 32-point sweep (pulse + 1000-point relaxation, '
                      'scipy LSODA rtol 1e-8 through the oracle port), %.1f s' % dt}


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port: `oracle/oracle_np.py` driving
    scipy LSODA) on all host cores.  One step = `cores` systems of the sweep (evenly spaced
    powers), one per core."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, min(len(os.sched_getaffinity(0)), 16))
    powers = powers_list()
    # bounded sample: one system per core and step, evenly spaced powers of the sweep
    idx = [int(round(i)) for i in np.linspace(0, N_POWERS - 1, cores)]
    jobs = [(2000, powers[i], True) for i in idx]
    # fresh interpreters (a forked copy of a process that already runs BLAS / OpenMP threads can
    # deadlock), one BLAS thread each so that `cores` means what it says
    for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[var] = '1'
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        for _ in range(min(args.warmup, 1)):           # CPU needs no warm-up beyond imports
            pool.map(_cpu_solve_one, [(2000, powers[0], True)] * 1)
        t0 = time.time()
        for _ in range(args.steps):
            pool.map(_cpu_solve_one, jobs)
        dt = time.time() - t0
    value = cores * args.steps / dt
    sample = ('%d of the 32 powers per step (indices %s), one per core, scipy LSODA rtol 1e-8 '
              'through the oracle port' % (cores, idx))
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
                        'rtol 1e-8 atol 1e-15; one lattice realisation per GPU',
            'systems_per_step_per_gpu': N_POWERS, 'n_gpus': n_gpus,
            'l2': 'L2 flushed (256 MiB write) between timed steps; the pair stream of this '
                  'lattice (1.2 MB) is L2-resident inside a solve by construction',
            'parallelism': 'independent systems, %d lattice realisation(s), no collective' % n_gpus}


# ------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from simetuc_b200 import _lib, lattice, precalculate
    from simetuc_b200.simulations import Simulations, cumulative_trapezoid

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

    # ---- setup (untimed): lattice realisation of this rank, device-resident system ------------
    cte = make_cte()
    powers = powers_list()
    lat = lattice.generate(cte, seed=2000 + rank)
    system = precalculate.build_system(cte, lat, device=local, n_sys=N_POWERS)
    for k, P in enumerate(powers):
        for exc in cte['excitations']['NIR_980']:
            exc.power_dens = P
        system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=k)
    N, kinds = system.N, system.sS + system.sA
    t_pulse, t_sol = t_grids(cte)
    _, _, pop = system.ground_states()
    ext = torch.cuda.ExternalStream(system.stream, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    d_tp = torch.tensor(t_pulse, **f64)
    d_ts = torch.tensor(t_sol, **f64)
    d_y0 = torch.empty(N_POWERS, N, **f64)
    system.to_planes(np.tile(pop.astype(np.float64), (N_POWERS, 1)), d_y0.data_ptr())
    d_yp = torch.empty(N_POWERS, 2, N, **f64)
    d_ys = torch.empty(N_POWERS, 1000, N, **f64)
    d_avg = torch.empty(N_POWERS, 1000, kinds, **f64)
    plane_of_ref = torch.tensor(np.argsort(_ref_of_plane(system)), device=dev)   # ref index -> plane
    ref_of_plane = torch.tensor(_ref_of_plane(system), device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, **f64)
    d_dt = torch.tensor(np.diff(t_sol - t_sol[0]), **f64)[None, :, None]
    kernel_ms, passes, stats_last = [], [], None

    def step_device():
        """One pass of the hot path with every operand resident in HBM."""
        st_p = system.integrate_dev(d_tp.data_ptr(), 2, d_y0.data_ptr(), N_POWERS, True, RTOL, ATOL,
                                    y_out_ptr=d_yp.data_ptr())
        ms_p, _ = system.last_timing()
        with torch.cuda.stream(ext):
            y0_relax = d_yp[:, -1, :].index_select(1, ref_of_plane).contiguous()
        st_r = system.integrate_dev(d_ts.data_ptr(), 1000, y0_relax.data_ptr(), N_POWERS, False,
                                    RTOL, ATOL, y_out_ptr=d_ys.data_ptr(), avg_out_ptr=d_avg.data_ptr())
        ms_r, _ = system.last_timing()
        with torch.cuda.stream(ext):
            # pulsed steady state: area under the emission curves (simulations.py:579-591)
            steady = (0.5 * (d_avg[:, 1:, :] + d_avg[:, :-1, :]) * d_dt).sum(dim=1)
        return ms_p + ms_r, st_p, st_r, steady

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if os.environ.get('BENCH_NO_SAMPLER'):          # diagnosis only: does NVML polling cost time?
        sampler.nv = None
    sampler.start()
    launches0 = system.lib.stb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record(ext)
    for _ in range(args.steps):
        with torch.cuda.stream(ext):
            flush.zero_()
        ms, st_p, st_r, steady = step_device()
        kernel_ms.append(ms)
        passes.append((sum(s['n_rhs'] for s in st_p + st_r), sum(s['n_jv'] for s in st_p + st_r)))
        stats_last = (st_p, st_r)
    ev1.record(ext)
    barrier()
    wall = time.time() - t0
    dev_ms = ev0.elapsed_time(ev1)
    launches = system.lib.stb_launch_count() - launches0
    clocks = sampler.stop()
    ok = all(s['status'] == 0 for s in stats_last[0] + stats_last[1])

    # ---- end to end through the public API with host buffers ----------------------------------
    sim = Simulations(make_cte(), devices=[local])
    sim.cte['lattice_seed'] = 2000 + rank
    precalculate._LATTICE_CACHE[precalculate._lattice_key(sim.cte)] = lat
    sim.simulate_power_dependence(powers, keep_y=False)             # warm-up
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t1 = time.time()
    for _ in range(e2e_steps):
        sol = sim.simulate_power_dependence(powers, keep_y=False)
        intens = sol.populations                                   # host result of the step
    torch.cuda.synchronize()
    e2e_wall = time.time() - t1
    h2d = (lat.positions.nbytes + lat.ion_type.nbytes + 72 + N_POWERS * 2 * 8 * (4 + 81) * 2 +
           2 * N_POWERS * N * 8 + (2 + 1000) * 8)
    d2h = N_POWERS * 2 * N * 8 + N_POWERS * 1000 * kinds * 8 + 2 * N_POWERS * 88 + 2 * system.n_ions * 8 * 2
    # the device-resident and the end-to-end path must agree on the emission intensities
    conc = np.array([5.0] * system.sS + [0.5] * system.sA)
    dev_int = steady.cpu().numpy() * conc
    agree = float(np.max(np.abs(dev_int - intens) / (np.abs(intens) + 1e-300)))

    # ---- max over ranks ----------------------------------------------------------------------------
    times = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max, e2e_ms_max = [float(v) for v in times.tolist()]
    n_rhs = float(np.mean([p[0] for p in passes]))
    n_jv = float(np.mean([p[1] for p in passes]))
    bpp = system.bytes_per_pair                # 8: dense all-to-all blocks (f64 d^-6 per ordered pair)
    b_rhs = bpp * system.n_pairs + 16 * N      # + y read, dy written once
    b_jv = bpp * system.n_pairs + 24 * N
    kms = float(np.mean(kernel_ms))
    achieved = (n_rhs * b_rhs + n_jv * b_jv) / (kms * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = N_POWERS * world * args.steps / (dev_ms_max * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'solves/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': dev_ms_max / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(world), 'clocks': clocks,
        'e2e': {'value': N_POWERS * world * e2e_steps / (e2e_ms_max * 1e-3), 'unit': 'solves/s',
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'api': 'Simulations.simulate_power_dependence(powers, keep_y=False): lattice -> '
                       'device pair build -> pulse + relaxation -> emission intensities on the host',
                'steps': e2e_steps, 'max_rel_diff_vs_device_path': agree},
        'gpu_launches': int(launches),
        'wall_ms_per_step': wall_ms_max / args.steps,
        'solver': {'ok': bool(ok), 'rhs_passes_per_step': n_rhs, 'jv_passes_per_step': n_jv,
                   'steps_min_max': [min(s['n_steps'] for s in stats_last[1]),
                                     max(s['n_steps'] for s in stats_last[1])]},
        'roofline': {'bound': 'hbm', 'kernel': 'k_integrate_cl_t<false> (cluster BDF/Newton-Krylov integrator; its '
                     'inner loop is the pair-stream pass, pair rows resident in shared memory)',
                     'achieved': achieved, 'peak': peaks['hbm_gbs'],
                     'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'], 'traffic': 2.2e6,
                     'traffic_source': 'profiles/r01_ncu_integrate_cl.txt: dram read+write of one '
                                       'launch of this batch with 40 output times (pair rows, state '
                                       'and Krylov vectors live in shared memory; the full run adds '
                                       'its trajectory output, 181 MB per launch)',
                     'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': n_rhs * b_rhs + n_jv * b_jv,
                     'note': 'latency-bound regime: the 89,700-pair lattice (0.7 MB) is resident in the '
                             'shared memory of its cluster, algorithmic bytes are a unit of account; '
                             'the HBM-bound numbers are in roofline_kernels'},
    }
    if not args.no_large:
        line['roofline_kernels'] = large_lattice_roofline(local, peaks, peak_src)
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


def large_lattice_roofline(device, peaks, peak_src):
    """RHS and Jv kernels on an HBM-bound lattice (SURVEY.md 8d "C4" shape: N_uc=40, 25 % Yb,
    0.3 % Tm, all ordered pairs -> 5.8e8 pairs, 7 GB of pair data, >> 126 MB of L2), timed with
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
                     'traffic': 4.71e9,
                     'traffic_source': 'profiles/r01_ncu_rhs_jv_tma.txt (ncu --set full of k_op_tma: dram read+write per launch)'}
    res['workload'] = 'N_uc=40 25%%Yb 0.3%%Tm: %d ions, %d states, %d ordered pairs, %.2f GB pair stream; ' \
                      'inputs >> L2' % (system.n_ions, N, system.n_pairs, system.stream_bytes / 1e9)
    res['peak_source'] = peak_src
    system.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-large', action='store_true', help='skip the HBM-bound RHS/Jv measurement')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline sample')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
