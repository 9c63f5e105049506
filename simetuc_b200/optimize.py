"""The optimisation objective - host mirror of `simetuc/optimize.py:79-133`.

The reference evaluates one candidate parameter set per objective call, sequentially, and each
evaluation rebuilds every matrix (`optimize.py:84-91`).  A candidate only rescales the rates of
whole processes (`w = strength * d^-mult`, `precalculate.py:289`) or changes a branching ratio,
so the lattice geometry is shared: `objective_batch` evaluates any number of candidates as one
batch of ODE systems on the device(s) (`stb_set_strengths` / `stb_set_blocks` per system).
The `lmfit` driver loop (`optimize.py:182-258`) is host glue outside this path; any driver that
can propose several candidates per iteration (brute-force grids, finite-difference gradients,
population methods) can call `objective_batch` directly.
"""
import copy
import functools
from typing import Callable, List, Sequence

import numpy as np

from . import simulations


def _values(params) -> List[float]:
    """Parameter values in order: an `lmfit.Parameters`-like object (`.valuesdict()`), a dict, or
    a plain sequence."""
    if hasattr(params, 'valuesdict'):
        return list(params.valuesdict().values())
    if isinstance(params, dict):
        return [getattr(v, 'value', v) for v in params.values()]
    return list(params)


def optim_fun(function: Callable, params, sim: simulations.Simulations) -> np.ndarray:
    '''Update parameter values, simulate and return the per-state error vector
    (`optimize.py:79-107`).  `function` is e.g. `sim.simulate_dynamics` with no parameters and
    must return an object with an `errors` attribute.'''
    for num, new_value in enumerate(_values(params)):
        sim.cte['optimization']['processes'][num].value = new_value
    if not sim.cte['optimization'].get('excitations', False):
        return function().errors
    for exc_lst in sim.cte['excitations'].values():
        for excitation in exc_lst:
            excitation.active = False
    n_kinds = sim.cte['states']['activator_states'] + sim.cte['states']['sensitizer_states']
    total_errors = np.zeros((n_kinds,), dtype=np.float64)
    for exc_label in sim.cte['optimization']['excitations']:
        sim.cte['excitations'][exc_label][0].active = True
        solution = function()
        sim.cte['excitations'][exc_label][0].active = False
        total_errors += solution.errors ** 2
    return np.sqrt(total_errors)


def optim_fun_dynamics(params, sim: simulations.Simulations, average: bool = False,
                       N_samples: int = None) -> np.ndarray:
    '''`optimize.py:109-117`.'''
    if N_samples is None:
        function = functools.partial(sim.simulate_dynamics, average=average)
    else:
        function = functools.partial(sim.sample_simulation, sim.simulate_dynamics,
                                     N_samples=N_samples, average=average)
    return optim_fun(function, params, sim)


def optim_fun_dynamics_conc(params, sim: simulations.Simulations, average: bool = False,
                            N_samples: int = None) -> np.ndarray:
    '''`optimize.py:119-133`.'''
    cd = sim.cte['concentration_dependence']
    if N_samples is None:
        function = functools.partial(sim.simulate_concentration_dependence, cd['concentrations'],
                                     cd['N_uc_list'], dynamics=True, average=average)
    else:
        function = functools.partial(sim.sample_simulation, sim.simulate_concentration_dependence,
                                     N_samples=N_samples, concentrations=cd['concentrations'],
                                     N_uc_list=cd['N_uc_list'], dynamics=True, average=average)
    return optim_fun(function, params, sim)


def objective_batch(sim: simulations.Simulations, candidates: Sequence[Sequence[float]],
                    keep_y: bool = False, on_device: bool = False):
    '''Evaluate the dynamics objective for many candidate parameter sets at once.

    `candidates[c][k]` is the value of `cte.optimization['processes'][k]` (an ET strength or a
    branching ratio / decay rate) for candidate c.  All candidates share the lattice and are
    dealt round-robin over `sim.devices`.  Returns `(errors [n_candidates, sS+sA], solutions)`,
    the per-state deviation from the experimental curves of every candidate exactly as
    `optim_fun` would return it one by one.

    `on_device=True`: interpolation onto the experimental time base and the per-state deviations
    (`simulations.py:425-527`) run on the GPU; the emission curves are not copied back and
    `solutions` is None (an optimiser only needs the errors).'''
    candidates = [list(map(float, c)) for c in candidates]
    shared = {key: sim.cte[key] for key in ('energy_transfer', 'decay', 'optimization')}
    ctes = []
    for values in candidates:
        # private copies of the sections a candidate changes, made in one deepcopy so that the
        # optimised objects listed in `optimization.processes` stay the ones stored in
        # `energy_transfer` / `decay`; every other section is shared (read-only)
        cte = copy.copy(sim.cte)
        cte.update(copy.deepcopy(shared))
        cte['states'] = copy.copy(sim.cte['states'])
        for proc, value in zip(cte['optimization']['processes'], values):
            proc.value = value
        ctes.append(cte)
    if on_device:
        template = simulations.DynamicsSolution(np.zeros(1), np.zeros((1, 1)), [0], [0], ctes[0],
                                                own_cte=True)
        experiments = template.list_exp_data
        sols, index_S_i, index_A_j, lat, stats = sim._sweep_same_lattice(
            ctes, dynamics=True, keep_y=False, experiments=experiments)
        return np.array([st['errors'] for st in stats]), None
    sols, index_S_i, index_A_j, lat, stats = sim._sweep_same_lattice(ctes, dynamics=True,
                                                                    keep_y=keep_y)
    solutions, errors = [], []
    exp_data = None
    for cte, (t_sol, y_sol, avg), st in zip(ctes, sols, stats):
        cte['ions'] = {'total': lat.n_ions, 'sensitizers': lat.n_S, 'activators': lat.n_A}
        cte['states']['energy_states'] = lat.n_states
        sol = simulations.DynamicsSolution(t_sol, y_sol if y_sol is not None else avg,
                                           index_S_i, index_A_j, cte, own_cte=True)
        sol.set_avg_data(avg)
        if exp_data is None:
            exp_data = sol.list_exp_data       # the same files for every candidate: read them once
        else:
            sol.__dict__['list_exp_data'] = exp_data
        sol.solver_stats = st
        solutions.append(sol)
        errors.append(sol.errors)
    return np.array(errors), solutions
