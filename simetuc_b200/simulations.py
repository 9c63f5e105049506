"""Orchestration and result objects - host mirror of `simetuc/simulations.py`.

`Simulations(cte)` keeps the reference's methods (`simulations.py:1045-1442`).  What changes
underneath: the sweeps that the reference runs as sequential Python loops, each one rebuilding
every matrix (`:1294-1305`, `:1356-1372`, `:1412-1417`), become batches of independent ODE
systems on the device: systems that share a lattice (powers) share one pair stream and are
integrated by one persistent kernel launch; with several GPUs the systems are dealt round-robin
(`system_id % n_gpus`), one host thread per device, and only populations / emission curves are
gathered at the end (no collective on the hot path).

The `Solution` family keeps the attributes the rest of simetuc reads (`t_sol, y_sol,
index_S_i, index_A_j, cte, list_avg_data, steady_state_populations, power_dens,
concentration, errors`); plotting and HDF5 persistence are outside this path (results are
saved as `.npz`).
"""
import copy
import csv
import functools
import logging
import os
import threading
import time
import warnings
from typing import Any, Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import odesolver, precalculate
from .util import Conc, disable_loggers


class cached_property:   # pylint: disable=invalid-name
    """Computes the value once and stores it on the instance (reference: `util.py:49-75`)."""

    def __init__(self, func):
        self.func = func
        self.__doc__ = func.__doc__

    def __get__(self, obj, cls):
        if obj is None:
            return self
        value = obj.__dict__[self.func.__name__] = self.func(obj)
        return value


def cumulative_trapezoid(y, x):
    """`scipy.integrate.cumtrapz(y, x, initial=0)` along axis 0 (removed from scipy >= 1.14,
    used by the reference at `simulations.py:583`)."""
    out = np.zeros_like(y)
    dt = np.diff(x).reshape((-1,) + (1,) * (y.ndim - 1))
    out[1:] = np.cumsum(0.5 * (y[1:] + y[:-1]) * dt, axis=0)
    return out


class Solution:
    '''Base class for solutions of rate equation problems (`simulations.py:41-286`).'''

    def __init__(self, t_sol: np.array, y_sol: np.array, index_S_i: List[int],
                 index_A_j: List[int], cte, average: bool = False, own_cte: bool = False) -> None:
        self.t_sol = t_sol
        self.y_sol = y_sol
        # the reference copies the settings (`simulations.py:64`); `own_cte=True`: the caller
        # hands over a private copy it made for this solution (the sweeps do)
        self.cte = cte if own_cte else copy.deepcopy(cte)
        self.index_S_i = list(index_S_i)
        self.index_A_j = list(index_A_j)
        self.average = average
        self.time = 0.0
        self.cte['colors'] = 'bk' if average else 'rk'
        self._prefix = 'solution'

    def __bool__(self) -> bool:
        return (self.t_sol.size != 0 and self.y_sol.size != 0 and bool(self.cte) and
                len(self.index_S_i) != 0 and len(self.index_A_j) != 0)

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, Solution):
            return NotImplemented
        return (self.y_sol.shape == other.y_sol.shape and np.allclose(self.t_sol, other.t_sol) and
                np.allclose(self.y_sol, other.y_sol) and
                self.index_S_i == other.index_S_i and self.index_A_j == other.index_A_j)

    def __ne__(self, other: object) -> bool:
        eq = self.__eq__(other)
        return eq if eq is NotImplemented else not eq

    def __repr__(self) -> str:
        return '{}(num_states={}, {}, power_dens={:.1e})'.format(
            self.__class__.__name__, self.y_sol.shape[1], self.concentration, self.power_dens)

    def _calculate_avg_populations(self) -> List[np.array]:
        '''Average population of each state kind over the ions of its type, clipped at zero;
        first S then A states (`simulations.py:100-128`).'''
        nS, nA = self.cte['states']['sensitizer_states'], self.cte['states']['activator_states']
        n_t = self.y_sol.shape[0]
        out = []
        for index, n in ((np.asarray(self.index_S_i), nS), (np.asarray(self.index_A_j), nA)):
            gs = index[index != -1]
            for state in range(n):
                if len(gs):
                    out.append(np.clip(self.y_sol[:, gs + state].sum(axis=1) / len(gs), 0, None))
                else:
                    out.append(np.zeros((n_t,)))
        return out

    @cached_property
    def list_avg_data(self) -> List[np.array]:
        '''List of average populations for each state in the solution'''
        matrix = getattr(self, '_avg_matrix', None)
        if matrix is not None:
            return [np.ascontiguousarray(col) for col in matrix.T]
        return self._calculate_avg_populations()

    def set_avg_matrix(self, avg: np.ndarray) -> None:
        '''Like `set_avg_data`, but the per-state curves are only split out of the `[n_t, sS+sA]`
        matrix when `list_avg_data` is read (a sweep over many systems mostly needs the last row).'''
        self._avg_matrix = avg
        self.__dict__.pop('list_avg_data', None)

    def set_avg_data(self, avg: np.ndarray) -> None:
        '''Use the per-state averages computed on the device ([n_t, sS+sA]) instead of the host
        loop above.'''
        self.__dict__['list_avg_data'] = [np.ascontiguousarray(col) for col in avg.T]

    @cached_property
    def state_labels(self) -> List[str]:
        st = self.cte['states']
        return ([st['sensitizer_ion_label'] + '_' + s for s in st['sensitizer_states_labels']] +
                [st['activator_ion_label'] + '_' + s for s in st['activator_states_labels']])

    @cached_property
    def power_dens(self) -> float:
        '''Return the power density used to obtain this solution (first active excitation).'''
        for excitation in self.cte['excitations'].values():
            if excitation[0].active:
                return excitation[0].power_dens
        raise AttributeError('This Solution has no active excitation!')

    @cached_property
    def concentration(self) -> Conc:
        return Conc(self.cte['lattice']['S_conc'], self.cte['lattice']['A_conc'])

    def save(self, full_path: str) -> None:
        '''Save t_sol, y_sol, averages and ground-state indices (`.npz`; the reference writes the
        same members to HDF5, `simulations.py:227-241`).'''
        np.savez_compressed(full_path, t_sol=self.t_sol, y_sol=self.y_sol,
                            y_sol_avg=np.array(self.list_avg_data),
                            index_S_i=np.array(self.index_S_i), index_A_j=np.array(self.index_A_j))

    def save_txt(self, full_path: str, mode: str = 'wt') -> None:
        '''Average populations as text, one column per state (`simulations.py:243-263`).'''
        with open(full_path, mode) as fh:
            fh.write('#t ' + ' '.join(self.state_labels) + '\n')
            np.savetxt(fh, np.column_stack([self.t_sol] + list(self.list_avg_data)), fmt='%1.4e')


class SteadyStateSolution(Solution):
    '''Solution to a steady state problem (`simulations.py:288-327`).'''

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._prefix = 'steady_state'

    def _calculate_final_populations(self) -> np.ndarray:
        '''Last point of every average curve times the ion concentration
        (`simulations.py:302-312`).'''
        st, lat = self.cte['states'], self.cte['lattice']
        conc = np.array([lat['S_conc']] * st['sensitizer_states'] +
                        [lat['A_conc']] * st['activator_states'], dtype=np.float64)
        matrix = getattr(self, '_avg_matrix', None)
        last = matrix[-1] if matrix is not None else np.array([curve[-1] for curve in self.list_avg_data])
        return last * conc

    @cached_property
    def steady_state_populations(self) -> np.ndarray:
        return self._calculate_final_populations()


class DynamicsSolution(Solution):
    '''Solution to a dynamics problem; loads experimental decay curves and computes the
    per-state deviation (`simulations.py:330-601`).'''

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._prefix = 'dynamics'

    @staticmethod
    def _load_exp_data(filename: str, lattice_name: str, filter_window: int = 35):
        '''Two-column decay curve from `expData/<lattice>/<filename>`, normalised to its smoothed
        maximum (`simulations.py:354-422`).  Returns None when the file does not exist.'''
        from scipy import signal
        root = 'expData' if os.path.isdir('expData') else os.path.join('simetuc', 'expData')
        path = os.path.join(root, lattice_name, filename)
        if not os.path.isfile(path):
            return None
        data = None
        for delim in (' ', '\t', ',', ';'):
            with open(path, 'rt') as fh:
                rows = [row for row in csv.reader(fh, delimiter=delim)
                        if len(row) == 2 and not row[0].startswith('#')]
            try:
                data = np.array(rows, dtype=np.float64)
            except ValueError:
                continue
            if data.ndim == 2 and data.shape[1] == 2:
                break
            data = None
        if data is None:
            warnings.warn('Invalid experimental data in file "{}", it will be ignored.'.format(filename))
            return None
        smooth = signal.savgol_filter(data[:, 1], filter_window, 5, mode='nearest').clip(min=0)
        peak = int(np.argmax(smooth))
        delta = max(int(0.005 * len(smooth)), 1)
        if peak == 0:
            avg_max = smooth.max()
        elif peak > delta // 2:
            avg_max = np.mean(smooth[peak - delta // 2: peak + delta // 2])
        else:
            avg_max = np.mean(smooth[:peak + delta // 2])
        data[:, 1] = (data[:, 1] - smooth.min()) / (avg_max - smooth.min())
        return data.clip(min=0)

    @staticmethod
    def _get_background(exp_data, sim_data, offset_points: int = 50):
        '''Experimental background scaled to the simulated maximum (`simulations.py:425-449`).'''
        if exp_data is None or not np.any(exp_data) or not np.any(sim_data):
            return 0
        counts = exp_data[:, 1]
        tail = counts[counts > 0][-offset_points:]
        offset = np.mean(tail) * np.max(sim_data) if tail.size else 0
        return offset if np.isfinite(offset) and offset > 0 else 0

    @cached_property
    def list_exp_data(self) -> List[Optional[np.ndarray]]:
        '''Experimental decay curves named after state, excitation and concentrations
        (`simulations.py:529-549`).'''
        active = [label for label, exc in self.cte['excitations'].items() if exc[0].active]
        lat, st = self.cte['lattice'], self.cte['states']
        conc = '_{}{}_{}{}'.format(float(lat['S_conc']), st['sensitizer_ion_label'],
                                   float(lat['A_conc']), st['activator_ion_label'])
        return [self._load_exp_data('decay_' + label + '_exc_' + '_'.join(active) + conc + '.txt',
                                    lat['name']) for label in self.state_labels]

    @cached_property
    def list_avg_data_ofs(self) -> List[np.ndarray]:
        '''Average curves with the experimental background added (`simulations.py:452-462`).'''
        return [sim + self._get_background(exp, sim)
                for exp, sim in zip(self.list_exp_data, self.list_avg_data)]

    @cached_property
    def list_interp_data(self) -> List[np.ndarray]:
        '''Simulated curves on the experimental time base (`simulations.py:465-487`).'''
        out = []
        for exp, sim in zip(self.list_exp_data, self.list_avg_data):
            if exp is None or not np.any(exp):
                out.append(sim)
                continue
            offset = self._get_background(exp, sim)
            out.append(np.interp(exp[:, 0], self.t_sol, sim, left=offset, right=offset))
        return out

    def _calc_errors(self) -> np.ndarray:
        '''`1/mean(sim) * sqrt(mean((sim - exp*max(sim))^2))` per state, 0 where there is no
        experimental curve (`simulations.py:499-527`).'''
        errors = []
        for sim, exp in zip(self.list_interp_data, self.list_exp_data):
            if exp is None or sim is None:
                errors.append(0.0)
            else:
                dev = (sim - exp[:, 1] * np.max(sim)) ** 2
                errors.append(1 / np.mean(sim) * np.sqrt(np.sum(dev) / len(sim)))
        return np.array(errors)

    @cached_property
    def errors(self) -> np.ndarray:
        return self._calc_errors()

    @cached_property
    def total_error(self) -> float:
        return float(np.sqrt(np.sum(self.errors ** 2)))

    def calculate_steady_state(self) -> SteadyStateSolution:
        '''Steady state of a pulsed source: area under every decay curve
        (`simulations.py:579-591`).'''
        y_steady = cumulative_trapezoid(self.y_sol, self.t_sol)
        steady_sol = SteadyStateSolution(self.t_sol, y_steady, self.index_S_i, self.index_A_j,
                                         self.cte, average=self.average)
        steady_sol.time = self.time
        return steady_sol


class SolutionList(Sequence):
    '''Base class for a list of solutions (`simulations.py:646-760`).'''

    def __init__(self) -> None:
        self.solution_list = []  # type: List[Solution]
        self.time = 0.0
        self.dynamics = False
        self._prefix = 'solutionlist'

    def __bool__(self) -> bool:
        return len(self.solution_list) != 0

    def __len__(self) -> int:
        return len(self.solution_list)

    def __getitem__(self, index):
        return self.solution_list[index]

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, SolutionList):
            return NotImplemented
        return self.solution_list == other.solution_list

    def add_solutions(self, sol_list: List[Solution]) -> None:
        self.solution_list = list(sol_list)

    def save(self, full_path: str) -> None:
        for num, sol in enumerate(self.solution_list):
            sol.save('{}_{}'.format(full_path, num))


class PowerDependenceSolution(SolutionList):
    '''Solution to a power dependence simulation (`simulations.py:763-835`).'''

    def __init__(self) -> None:
        super().__init__()
        self._prefix = 'pow_dep'

    @property
    def power_dens_list(self) -> List[float]:
        return [sol.power_dens for sol in self.solution_list]

    @property
    def populations(self) -> np.ndarray:
        '''`[n_powers, sS+sA]` steady-state populations (emission intensities).'''
        return np.array([sol.steady_state_populations for sol in self.solution_list])


class ConcentrationDependenceSolution(SolutionList):
    '''Solution to a concentration dependence simulation (`simulations.py:838-1042`).'''

    def __init__(self, dynamics: bool = False) -> None:
        super().__init__()
        self.dynamics = dynamics
        self._prefix = 'conc_dep'

    @property
    def concentrations(self) -> List[Conc]:
        return [sol.concentration for sol in self.solution_list]


def _run_on_devices(jobs: List[Callable[[], Any]], n_threads: int) -> List[Any]:
    """Run independent jobs (each bound to one device) on `n_threads` host threads; the C ABI
    releases the GIL for the whole solve."""
    results: List[Any] = [None] * len(jobs)
    errors: List[BaseException] = []

    def worker(slot):
        for k in range(slot, len(jobs), n_threads):
            try:
                results[k] = jobs[k]()
            except BaseException as err:   # pylint: disable=broad-except
                errors.append(err)

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errors:
        raise errors[0]
    return results


class Simulations:
    '''Setup and solve a dynamics or a steady state problem (`simulations.py:1045-1442`).

    `devices`: CUDA device ordinals to spread independent systems over (default: device 0).'''

    def __init__(self, cte, full_path: str = None, devices: Optional[Sequence[int]] = None,
                 streams_per_device: int = 16) -> None:
        self.cte = cte
        self.full_path = full_path
        #: independent lattices (concentration points) solved concurrently on one device
        self.streams_per_device = max(1, int(streams_per_device))
        self.devices = list(devices) if devices else [0]

    def __bool__(self) -> bool:
        return bool(self.cte)

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, Simulations):
            return NotImplemented
        return self.cte == other.cte and self.full_path == other.full_path

    def __repr__(self) -> str:
        return '{}(lattice={}, n_uc={}, num_states={})'.format(
            self.__class__.__name__, self.cte['lattice']['name'], self.cte['lattice']['N_uc'],
            self.cte['states'].get('energy_states'))

    # ------------------------------------------------------------------ time grids
    def _get_t_pulse(self) -> float:
        '''Pulse width of the first active excitation (`simulations.py:1078-1090`).'''
        for excitation in self.cte['excitations'].values():
            if excitation[0].active:
                tf_p = excitation[0].t_pulse
                if tf_p is None:
                    break
                return tf_p
        raise ValueError('t_pulse value not found! Please add t_pulse to your excitation settings.')

    def _get_t_simulation(self) -> float:
        '''`round(15 * max lifetime, 8)` (`simulations.py:1092-1093`).'''
        return float(np.round(15 * np.max(precalculate.get_lifetimes(self.cte)), 8))

    def _is_pulsed(self) -> bool:
        return any(exc.active and exc.t_pulse is not None
                   for lst in self.cte['excitations'].values() for exc in lst)

    # ------------------------------------------------------------------ single systems
    def simulate_dynamics(self, average: bool = False) -> DynamicsSolution:
        '''Pulse then relaxation on 1000 log-spaced points (`simulations.py:1096-1176`);
        `average=True`: the average rate equations (`:1114-1119`).'''
        logger = logging.getLogger(__name__ + '.dynamics')
        start_time = time.time()
        gen_lattice = self.cte.get('gen_lattice', False)
        setup_func = precalculate.setup_average_eqs if average else precalculate.setup_microscopic_eqs
        (cte_updated, initial_population, index_S_i, index_A_j, total_abs_matrix, decay_matrix,
         ET_matrix, N_indices, jac_indices, coop_ET_matrix, coop_N_indices,
         coop_jac_indices) = setup_func(
             self.cte, full_path=self.full_path, gen_lattice=gen_lattice, device=self.devices[0])
        self.cte = cte_updated
        tf = self._get_t_simulation()
        tf_p = self._get_t_pulse()
        N_steps = self.cte['simulation_params']['N_steps']
        rtol = self.cte['simulation_params']['rtol']
        atol = self.cte['simulation_params']['atol']
        t_pulse = np.linspace(0, tf_p, 2, dtype=np.float64)
        y_pulse = odesolver.solve_pulse(t_pulse, initial_population.transpose(), total_abs_matrix,
                                        decay_matrix, ET_matrix, N_indices, jac_indices,
                                        coop_ET_matrix, coop_N_indices, coop_jac_indices,
                                        rtol=rtol, atol=atol, quiet=self.cte['no_console'])
        t_sol = np.logspace(np.log10(tf_p), np.log10(tf), N_steps, dtype=np.float64)
        y_sol = odesolver.solve_relax(t_sol, y_pulse[-1, :], decay_matrix, ET_matrix, N_indices,
                                      jac_indices, coop_ET_matrix, coop_N_indices,
                                      coop_jac_indices, rtol=rtol, atol=atol,
                                      quiet=self.cte['no_console'])
        t_sol = t_sol - t_sol[0]
        dynamics_sol = DynamicsSolution(t_sol, y_sol, index_S_i, index_A_j, self.cte, average=average)
        dynamics_sol.time = time.time() - start_time
        logger.info('Simulation finished! Total time: %.2fs.', dynamics_sol.time)
        return dynamics_sol

    def simulate_steady_state(self, average: bool = False) -> SteadyStateSolution:
        '''Pulsed source -> area under the decay curves, CW source -> long-time limit
        (`simulations.py:1184-1193`).'''
        if self._is_pulsed():
            return self.simulate_pulsed_steady_state(average=average)
        return self.simulate_CW_steady_state(average=average)

    def simulate_CW_steady_state(self, average: bool = False) -> SteadyStateSolution:
        '''`linspace(0, 15 tau_max, N_steps)` under constant excitation (`simulations.py:1195-1257`).'''
        start_time = time.time()
        setup_func = precalculate.setup_average_eqs if average else precalculate.setup_microscopic_eqs
        (cte_updated, initial_population, index_S_i, index_A_j, total_abs_matrix, decay_matrix,
         ET_matrix, N_indices, jac_indices, coop_ET_matrix, coop_N_indices,
         coop_jac_indices) = setup_func(
             self.cte, full_path=self.full_path, device=self.devices[0])
        self.cte = cte_updated
        tf = self._get_t_simulation()
        N_steps = self.cte['simulation_params']['N_steps']
        rtol = self.cte['simulation_params']['rtol']
        atol = self.cte['simulation_params']['atol']
        t_pulse = np.linspace(0, tf, N_steps)
        y_pulse = odesolver.solve_pulse(t_pulse, initial_population.transpose(), total_abs_matrix,
                                        decay_matrix, ET_matrix, N_indices, jac_indices,
                                        coop_ET_matrix, coop_N_indices, coop_jac_indices,
                                        nsteps=1000, method='bdf', rtol=rtol, atol=atol,
                                        quiet=self.cte['no_console'])
        steady_sol = SteadyStateSolution(t_pulse, y_pulse, index_S_i, index_A_j, self.cte,
                                         average=average)
        steady_sol.time = time.time() - start_time
        return steady_sol

    def simulate_pulsed_steady_state(self, average: bool = False) -> SteadyStateSolution:
        '''`simulations.py:1265-1269`.'''
        return self.simulate_dynamics(average=average).calculate_steady_state()

    # ------------------------------------------------------------------ batched sweeps
    def _solve_batch(self, lat, ctes: List[Any], device: int, dynamics: bool,
                     strengths: Optional[np.ndarray] = None, keep_y: bool = True,
                     average: bool = False, gpu_busy: Optional[threading.Event] = None,
                     sm_budget: Optional[int] = None, experiments: Optional[list] = None,
                     steady: bool = False):
        """Integrate `len(ctes)` systems that share `lat` in one batch on `device`.
        Each cte supplies its own absorption / decay blocks (and optionally ET strengths).
        Returns a list of `(t_sol, y_sol | None, avg)` per system plus the solver statistics.

        The device system of the lattice is checked out of `precalculate`'s pool (pair build and
        device buffers are reused by the next sweep on this lattice); pulse and relaxation run
        in one C-ABI call.  `experiments` (dynamics only): one experimental curve or None per
        state kind; the per-state deviations are then evaluated on the device and returned as
        `stats[k]['errors']`, and the curves are not copied back (`avg` is None).  `steady`
        (dynamics only): `avg` is the running time integral of the emission curves, computed on the
        device (the pulsed steady state of `DynamicsSolution.calculate_steady_state`).
        Systems whose solve did not reach the end are reported with the
        reference's warnings (`odesolver.py:129-134`); a relaxation that would start from an
        unreached pulse row is blanked (NaN) instead of returned as data."""
        logger = logging.getLogger(__name__)
        if average:       # same systems, processes at their average strengths (`precalculate.py:999-1003`)
            ctes = [precalculate.average_view(cte) for cte in ctes]
        cte0 = ctes[0]
        n = len(ctes)
        coop0 = [(name, p.strength) for name, p in cte0['energy_transfer'].items() if len(p.type) == 3]
        for cte in ctes[1:]:
            if [(name, p.strength) for name, p in cte['energy_transfer'].items()
                    if len(p.type) == 3] != coop0:
                raise ValueError('systems of one batch must share the cooperative process strengths '
                                 '(they are part of the device-resident lattice data)')
        system = precalculate.acquire_system(cte0, lat, device=device, n_sys=n)
        try:
            # share of the GPU of this solve when several lattices are in flight on the device
            system.set_option('sm_budget', sm_budget if sm_budget else 1 << 20)
            blocks = {}
            for k, cte in enumerate(ctes):
                # the decay blocks depend on the decay section only: sweeps share that object
                dkey = id(cte['decay'])
                if dkey not in blocks:
                    blocks[dkey] = precalculate.decay_blocks(cte)
                system.set_decay_blocks(*blocks[dkey], system=k)
                system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=k)
                if strengths is not None:
                    system.set_strengths(strengths[k], system=k)
                else:
                    system.set_strengths([p.strength for p in cte['energy_transfer'].values()
                                          if len(p.type) == 2] or np.zeros(0), system=k)
            index_S_i, index_A_j, pop = system.ground_states()
            y0 = np.tile(pop.astype(np.float64), (n, 1))
            params = cte0['simulation_params']
            rtol, atol, N_steps = params['rtol'], params['atol'], params['N_steps']
            tf = self._get_t_simulation()
            if gpu_busy is not None:
                gpu_busy.set()      # the long call: the GIL is free until it returns
            if dynamics:
                tf_p = self._get_t_pulse()
                t_pulse = np.linspace(0, tf_p, 2, dtype=np.float64)
                t_sol = np.logspace(np.log10(tf_p), np.log10(tf), N_steps, dtype=np.float64)
                errors = None
                if experiments is not None:
                    for kind, exp in enumerate(experiments):
                        system.set_experiment(kind, exp)
                _, y, avg, st_p, st = system.integrate_dynamics(
                    t_pulse, t_sol, y0, rtol=rtol, atol=atol, want_y=keep_y,
                    want_avg=experiments is None, cumulative=steady and experiments is None)
                if experiments is not None:
                    errors = system.objective(n)
                t_out = t_sol - t_sol[0]
                stats = [dict(pulse=a, relax=b) for a, b in zip(st_p, st)]
                for k, (a, b) in enumerate(zip(st_p, st)):
                    if errors is not None:
                        stats[k]['errors'] = errors[k]
                    if odesolver.warn_solver_status(a, logger, 'system %d, pulse: ' % k):
                        logger.warning('system %d: the pulse did not reach its end, its relaxation '
                                       'is not returned', k)
                        if avg is not None:
                            avg[k] = np.nan
                        if errors is not None:
                            errors[k] = np.nan
                        if y is not None:
                            y[k] = np.nan
                    else:
                        odesolver.warn_solver_status(b, logger, 'system %d, relaxation: ' % k)
            else:
                t_out = np.linspace(0, tf, N_steps)
                y, avg, st = system.integrate(t_out, y0, pulse=True, rtol=rtol, atol=atol,
                                              nsteps=1000, want_y=keep_y, want_avg=True)
                stats = [dict(pulse=a) for a in st]
                for k, a in enumerate(st):
                    odesolver.warn_solver_status(a, logger, 'system %d: ' % k)
            return [(t_out, None if y is None else y[k], None if avg is None else avg[k])
                    for k in range(n)], list(index_S_i), list(index_A_j), stats
        finally:
            precalculate.release_system(system)

    def _sweep_same_lattice(self, ctes: List[Any], dynamics: bool, keep_y: bool = True,
                            strengths: Optional[np.ndarray] = None, average: bool = False,
                            gpu_busy: Optional[threading.Event] = None,
                            experiments: Optional[list] = None, steady: bool = False):
        """Systems sharing one lattice, dealt round-robin over `self.devices`."""
        gen_lattice = self.cte.get('gen_lattice', False)
        if average:
            lat = precalculate.average_lattice(self.cte)
        else:
            lat = precalculate.get_lattice(self.cte, gen_lattice=gen_lattice, full_path=self.full_path)
        n_dev = min(len(self.devices), len(ctes))
        shards = [list(range(d, len(ctes), n_dev)) for d in range(n_dev)]
        jobs = [(lambda d=d, ids=ids: self._solve_batch(
            lat, [ctes[i] for i in ids], self.devices[d], dynamics, keep_y=keep_y, average=average,
            gpu_busy=gpu_busy, experiments=experiments, steady=steady,
            strengths=None if strengths is None else strengths[ids])) for d, ids in enumerate(shards)]
        results = _run_on_devices(jobs, n_dev)
        out = [None] * len(ctes)
        stats = [None] * len(ctes)
        for ids, (sols, index_S_i, index_A_j, st) in zip(shards, results):
            for i, sol, s in zip(ids, sols, st):
                out[i] = sol
                stats[i] = s
        return out, index_S_i, index_A_j, lat, stats

    def simulate_power_dependence(self, power_dens_list: List[float],
                                  average: bool = False, keep_y: bool = True
                                  ) -> PowerDependenceSolution:
        '''Steady state at every power density (`simulations.py:1277-1316`).  All powers share
        the lattice, the decay blocks and the pair stream; only the absorption block differs
        (`precalculate.py:99,112,120`), so they are integrated as one batch per device.'''
        logger = logging.getLogger(__name__ + '.pow_dep')
        start_time = time.time()
        power_dens_list = [float(elem) for elem in list(power_dens_list)]
        pulsed = self._is_pulsed()
        # The device only needs the absorption block of every power: the batch is launched from
        # light settings (shared sections, private excitations)
        light = []
        for power_dens in power_dens_list:
            cte = copy.copy(self.cte)
            cte['excitations'] = type(self.cte['excitations'])(
                (label, [copy.copy(exc) for exc in exc_list])
                for label, exc_list in self.cte['excitations'].items())
            for exc_list in cte['excitations'].values():
                for exc in exc_list:
                    exc.power_dens = power_dens
            light.append(cte)
        # keep_y=False on a pulsed source: the ion average commutes with the time integral, so the area
        # under the emission curves (running trapezoid, computed on the device) is the steady state
        # without moving the [n_t, N] trajectories to the host
        on_device_steady = pulsed and not keep_y
        sols, index_S_i, index_A_j, lat, stats = self._sweep_same_lattice(
            light, dynamics=pulsed, keep_y=keep_y, average=average, steady=on_device_steady)
        # per-solution settings (the reference deep-copies everything, `simulations.py:1290`): only
        # what differs between the solutions or is written per solution is private - the
        # excitations (made for the batch above), the state and ion counts; the other sections are
        # shared, read-only views of this Simulations' settings
        ctes = []
        for cte in light:
            cte['states'] = copy.copy(self.cte['states'])
            ctes.append(cte)
        solutions = []
        ions = {'total': lat.n_ions, 'sensitizers': lat.n_S, 'activators': lat.n_A}
        for k, (cte, (t_sol, y_sol, avg), st) in enumerate(zip(ctes, sols, stats)):
            cte['ions'] = dict(ions)
            cte['states']['energy_states'] = lat.n_states
            if pulsed and y_sol is not None:
                sol = DynamicsSolution(t_sol, y_sol, index_S_i, index_A_j, cte,
                                       average=average, own_cte=True).calculate_steady_state()
            elif pulsed:
                sol = SteadyStateSolution(t_sol, avg, index_S_i, index_A_j, cte, average=average,
                                          own_cte=True)
                sol.set_avg_matrix(avg)                 # `avg` is the steady-state (integrated) curves
            else:
                sol = SteadyStateSolution(t_sol, y_sol if y_sol is not None else avg,
                                          index_S_i, index_A_j, cte, average=average, own_cte=True)
                sol.set_avg_data(avg)
            sol.solver_stats = st
            solutions.append(sol)
        # the reference leaves the last power in the shared settings (`:1298-1300`)
        for exc_list in self.cte['excitations'].values():
            for exc in exc_list:
                exc.power_dens = power_dens_list[-1]
        power_dep_solution = PowerDependenceSolution()
        power_dep_solution.add_solutions(solutions)
        power_dep_solution.time = time.time() - start_time
        logger.info('Power dependence curves finished! Total time: %.2fs.', power_dep_solution.time)
        return power_dep_solution

    def simulate_concentration_dependence(self, concentrations: List[Tuple[float, float]],
                                          N_uc_list: List[int] = None, dynamics: bool = False,
                                          average: bool = False, keep_y: bool = True
                                          ) -> ConcentrationDependenceSolution:
        '''One lattice per (S_conc, A_conc) point (`simulations.py:1318-1385`); the points are
        independent and are dealt round-robin over the devices.  `keep_y=False` (dynamics only)
        brings back the emission curves (`list_avg_data`, computed on the device) instead of the
        `[n_t, N]` trajectories: `y_sol` of the solutions then holds the curves.'''
        start_time = time.time()
        concentrations = [(float(a), float(b)) for a, b in list(concentrations)]
        if N_uc_list is None:
            N_uc_list = [self.cte['lattice']['N_uc']] * len(concentrations)
        N_uc_list = list(N_uc_list)
        if len(N_uc_list) < len(concentrations):
            N_uc_list.extend([N_uc_list[-1]] * (len(concentrations) - len(N_uc_list)))
        pulsed = self._is_pulsed()

        base_seed = self.cte.get('lattice_seed')

        def solve_point(k, device):
            cte = copy.deepcopy(self.cte)
            cte['lattice']['N_uc'] = N_uc_list[k]
            cte['lattice']['S_conc'], cte['lattice']['A_conc'] = concentrations[k]
            if base_seed is not None:
                # every point is its own lattice realisation, also when a concentration repeats
                cte['lattice_seed'] = base_seed + k
            sim = Simulations(cte, devices=[device])
            if average:
                lat = precalculate.average_lattice(cte)
            else:
                lat = precalculate.get_lattice(cte, gen_lattice=cte.get('gen_lattice', False))
            want_y = keep_y or not dynamics
            sols, index_S_i, index_A_j, st = sim._solve_batch(lat, [cte], device, dynamics or pulsed,
                                                              average=average, keep_y=want_y,
                                                              sm_budget=sm_budget)
            cte['ions'] = {'total': lat.n_ions, 'sensitizers': lat.n_S, 'activators': lat.n_A}
            cte['states']['energy_states'] = lat.n_states
            t_sol, y_sol, avg = sols[0]
            if dynamics:
                sol = DynamicsSolution(t_sol, y_sol if y_sol is not None else avg, index_S_i,
                                       index_A_j, cte, average=average, own_cte=True)
                sol.set_avg_data(avg)
            elif pulsed:
                sol = DynamicsSolution(t_sol, y_sol, index_S_i, index_A_j, cte,
                                       average=average).calculate_steady_state()
            else:
                sol = SteadyStateSolution(t_sol, y_sol, index_S_i, index_A_j, cte, average=average)
                sol.set_avg_data(avg)
            sol.solver_stats = st[0]
            return sol

        # every point is its own lattice, i.e. its own handle and stream: several points are in
        # flight per device (a system of this size occupies one cluster of <= 8 SMs out of 148)
        n_dev = min(len(self.devices), len(concentrations))
        per_dev = max(1, min(self.streams_per_device, -(-len(concentrations) // n_dev)))
        sm_budget = max(8, 148 // per_dev)
        jobs = [(lambda k=k: solve_point(k, self.devices[k % n_dev]))
                for k in range(len(concentrations))]
        with disable_loggers([__name__ + '.dynamics', __name__ + '.steady_state']):
            solutions = _run_on_devices(jobs, n_dev * per_dev)
        self.cte['lattice']['N_uc'] = N_uc_list[-1]
        self.cte['lattice']['S_conc'], self.cte['lattice']['A_conc'] = concentrations[-1]
        conc_dep_solution = ConcentrationDependenceSolution(dynamics=dynamics)
        conc_dep_solution.add_solutions(solutions)
        conc_dep_solution.time = time.time() - start_time
        return conc_dep_solution

    def sample_simulation(self, simulation_fun: Callable, N_samples: int, *args: Any,
                          **kwargs: Any):
        '''Repeat `simulation_fun` over `N_samples` fresh lattice realisations and average the
        emission curves (`simulations.py:1387-1442`).

        The reference runs the samples one after the other (`:1412-1417`); they are independent
        lattices, so here they are dealt round-robin over `self.devices` and run concurrently,
        `streams_per_device` at a time per GPU (every realisation has its own device handle and
        stream).  `simulation_fun` must be a method of this object (or a `functools.partial` of
        one): every sample runs it on a private copy of the settings with its own lattice seed
        (`lattice_seed + i` when a seed is set, fresh entropy otherwise).  The result is the last
        sample's solution carrying the averaged curves and errors, as in the reference.'''
        if not N_samples or N_samples <= 1:
            return simulation_fun(*args, **kwargs)
        self.cte['gen_lattice'] = True
        base_seed = self.cte.get('lattice_seed')
        if base_seed is None:
            seeds = [int(s.generate_state(1)[0]) for s in np.random.SeedSequence().spawn(N_samples)]
        else:
            seeds = [base_seed + i for i in range(N_samples)]
        inner, extra_args, extra_kw = simulation_fun, (), {}
        if isinstance(inner, functools.partial):
            inner, extra_args, extra_kw = inner.func, inner.args, dict(inner.keywords or {})
        if getattr(inner, '__self__', None) is not self:
            raise TypeError('sample_simulation needs a method of this Simulations object')
        name = inner.__name__
        n_dev = min(len(self.devices), N_samples)

        def run_sample(i):
            cte = copy.deepcopy(self.cte)
            cte['lattice_seed'] = seeds[i]
            sim = Simulations(cte, full_path=self.full_path, devices=[self.devices[i % n_dev]],
                              streams_per_device=self.streams_per_device)
            sol = getattr(sim, name)(*extra_args, *args, **extra_kw, **kwargs)
            if isinstance(sol, DynamicsSolution):       # evaluated in the worker: the GPU of the
                _ = sol.errors, sol.total_error, sol.list_avg_data_ofs   # next sample is busy meanwhile
            return sol

        per_dev = max(1, min(self.streams_per_device, -(-N_samples // n_dev)))
        jobs = [(lambda i=i: run_sample(i)) for i in range(N_samples)]
        sols = _run_on_devices(jobs, n_dev * per_dev)
        sol = sols[-1]
        if isinstance(sol, DynamicsSolution):
            acc = sum(np.array(s.list_avg_data_ofs) for s in sols)
            sol.__dict__['list_avg_data_ofs'] = list(acc / N_samples)
            sol.__dict__['total_error'] = float(np.mean([s.total_error for s in sols]))
            sol.__dict__['errors'] = np.mean([s.errors for s in sols], axis=0)
        return sol
