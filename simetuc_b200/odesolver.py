"""ODE core of the microscopic rate equations - host mirror of `simetuc/odesolver.py`.

Same names, argument order and return shapes as the reference (`odesolver.py:33-209`); the
arithmetic runs on the B200 through the C ABI:

    reference                                   here
    _rate_eq / _rate_eq_pulse  (:33-43, :65-71)  fused pair-stream RHS kernel (`stb_rhs`)
    _jac_rate_eq[_pulse]       (:46-62, :74-80)  Jacobian-vector kernel (`stb_jv`); the dense
                                                 Jacobian is never needed by the integrator
    _solve_ode                 (:85-137)         persistent BDF / Newton-Krylov kernel
                                                 (`stb_integrate`) instead of scipy's VODE
    solve_pulse / solve_relax  (:178-209)        unchanged plumbing, one `_solve_ode` call each

The sparse operands are the ones returned by `precalculate.setup_microscopic_eqs`; they carry
the handle of the device-resident system (`matrix.stb`).  Operands without a handle (foreign
scipy matrices) are rejected: there is no CPU fallback.
"""
import logging
from typing import Callable, Tuple

import numpy as np


def _system_of(*operands):
    """`(RateSystem, system index)` behind the operands.  Operand sets created by separate
    `setup_*_eqs` calls on one lattice share a device handle: the parameters this set was built
    from (`precalculate.ParamSet`) are re-applied first, so earlier operands do not silently
    integrate with a later call's power or strengths."""
    for op in operands:
        ref = getattr(op, 'stb', None)
        if ref is not None:
            params = getattr(op, 'stb_params', None)
            if params is not None:
                params.apply(ref[0], ref[1])
            return ref
    raise TypeError('operands do not come from simetuc_b200.precalculate.setup_microscopic_eqs '
                    '(no device system attached); this path has no CPU fallback')


def _check_coop(coop_N_indices, system=None):
    """Cooperative operands are accepted when they describe the process the device system was
    built with (they are then views exported from it); anything else has no device counterpart."""
    n = 0 if coop_N_indices is None else np.shape(coop_N_indices)[0]
    if system is not None and n != system.n_coop:
        raise TypeError('cooperative operands do not match the device system (%d interactions, '
                        'the system holds %d)' % (n, system.n_coop))


def _rate_eq(t, y, decay_matrix, UC_matrix, N_indices, coop_ET_matrix, coop_N_indices):
    '''Calculates the rhs of the ODE for the relaxation (reference: `odesolver.py:33-43`).'''
    system, k = _system_of(UC_matrix, decay_matrix)
    _check_coop(coop_N_indices, system)
    return system.rhs(np.asarray(y, dtype=np.float64), pulse=False, first=k)


def _rate_eq_pulse(t, y, abs_matrix, decay_matrix, UC_matrix, N_indices,
                   coop_ET_matrix, coop_N_indices):
    '''Calculates the rhs of the ODE for the excitation pulse (`odesolver.py:65-71`).'''
    system, k = _system_of(UC_matrix, decay_matrix, abs_matrix)
    _check_coop(coop_N_indices, system)
    return system.rhs(np.asarray(y, dtype=np.float64), pulse=True, first=k)


def _jac_vec(t, y, v, decay_matrix, UC_matrix, pulse=False, abs_matrix=None):
    '''`J(y) @ v` for the relaxation (pulse=False) or pulse Jacobian, matrix-free.'''
    system, k = _system_of(UC_matrix, decay_matrix, abs_matrix)
    return system.jv(np.asarray(y, dtype=np.float64), np.asarray(v, dtype=np.float64),
                     pulse=pulse, first=k)


def _dense_jacobian(system, k, y, pulse):
    N = system.N
    J = np.empty((N, N))
    eye = np.eye(N)
    y = np.asarray(y, dtype=np.float64)
    for col in range(N):                 # API compatibility only; the integrator never does this
        J[:, col] = system.jv(y, eye[col], pulse=pulse, first=k)
    return J


def _jac_rate_eq(t, y, decay_matrix, UC_matrix, jac_indices, coop_ET_matrix, coop_jac_indices):
    '''Dense Jacobian of the relaxation (`odesolver.py:46-62`), assembled from N device Jv
    products.  Kept for interface parity; the device integrator is Jacobian-free.'''
    system, k = _system_of(UC_matrix, decay_matrix)
    _check_coop(None if coop_jac_indices is None else np.asarray(coop_jac_indices)[::3], system)
    return _dense_jacobian(system, k, y, pulse=False)


def _jac_rate_eq_pulse(t, y, abs_matrix, decay_matrix, UC_matrix, jac_indices,
                       coop_ET_matrix, coop_jac_indices):
    '''Dense Jacobian of the excitation pulse (`odesolver.py:74-80`).'''
    system, k = _system_of(UC_matrix, decay_matrix, abs_matrix)
    _check_coop(None if coop_jac_indices is None else np.asarray(coop_jac_indices)[::3], system)
    return _dense_jacobian(system, k, y, pulse=True)


def warn_solver_status(st, logger=None, label: str = '') -> bool:
    '''The reference's messages for a solver that could not reach the end (`odesolver.py:129-134`);
    returns True when `st` (one `stb_stats` dict) reports a failure.  Non-convergence is not an
    error: the caller keeps what was computed (rows not reached are zero).'''
    if st['status'] == 0:
        return False
    logger = logger or logging.getLogger(__name__)
    reasons = {1: 'Excess work done on this call (increase "nsteps").',
               2: 'Step size underflow.', 3: 'Non-finite values in the right-hand side.'}
    logger.warning('%s%s', label, reasons.get(st['status'], 'solver failure'))
    logger.warning('Most likely the ode solver is taking too many steps.')
    logger.warning('Either change your settings or increase "nsteps".')
    logger.warning('The program will continue, but the accuracy of the ' +
                   'results cannot be guaranteed.')
    return True


def _solve_ode(t_arr: np.array, fun: Callable, fargs: Tuple, jfun: Callable, jargs: Tuple,
               initial_population: np.array, rtol: float = 1e-3, atol: float = 1e-15,
               nsteps: int = 1000, method: str = 'bdf', quiet: bool = True) -> np.array:
    '''Solve the ode for the times t_arr (reference: `odesolver.py:85-137`).  `fun` selects the
    problem (`_rate_eq_pulse` or `_rate_eq`), `fargs` are its operands.  Returns
    `y_arr [len(t_arr), N]`, row 0 = the initial population.  Like the reference, a solver that
    cannot reach the end logs a warning and returns what was computed (`:129-134`).'''
    logger = logging.getLogger(__name__)
    if method != 'bdf':
        raise ValueError("only method='bdf' is available on the device path")
    pulse = fun is _rate_eq_pulse
    system, k = _system_of(*fargs)
    _check_coop(fargs[-1], system)
    y0 = np.asarray(initial_population, dtype=np.float64).ravel()
    y_out, _, stats = system.integrate(np.asarray(t_arr, dtype=np.float64), y0, pulse=pulse,
                                       rtol=rtol, atol=atol, nsteps=nsteps, first=k)
    st = stats[0]
    warn_solver_status(st, logger)
    _solve_ode.last_stats = st
    return y_out[0]


def solve_pulse(t_pulse: np.array, initial_pop: np.array, abs_matrix, decay_matrix, UC_matrix,
                N_indices: np.array, jac_indices: np.array, coop_ET_matrix,
                coop_N_indices: np.array, coop_jac_indices: np.array,
                nsteps: int = 1000, rtol: float = 1e-3, atol: float = 1e-15,
                quiet: bool = False, method: str = 'bdf') -> np.array:
    '''Solve the response to an excitation pulse (reference: `odesolver.py:178-193`).'''
    import simetuc_b200.odesolver as this          # module lookup: the reference's mock seam
    return this._solve_ode(t_pulse, _rate_eq_pulse,
                           (abs_matrix, decay_matrix, UC_matrix, N_indices,
                            coop_ET_matrix, coop_N_indices),
                           _jac_rate_eq_pulse,
                           (abs_matrix, decay_matrix, UC_matrix, jac_indices,
                            coop_ET_matrix, coop_jac_indices),
                           initial_pop, method=method,
                           rtol=rtol, atol=atol, nsteps=nsteps, quiet=quiet)


def solve_relax(t_sol: np.array, initial_pop: np.array, decay_matrix, UC_matrix,
                N_indices: np.array, jac_indices: np.array, coop_ET_matrix,
                coop_N_indices: np.array, coop_jac_indices: np.array,
                nsteps: int = 1000, rtol: float = 1e-3, atol: float = 1e-15,
                quiet: bool = False) -> np.array:
    '''Solve the relaxation after a pulse (reference: `odesolver.py:196-209`).'''
    import simetuc_b200.odesolver as this
    return this._solve_ode(t_sol, _rate_eq,
                           (decay_matrix, UC_matrix, N_indices, coop_ET_matrix, coop_N_indices),
                           _jac_rate_eq,
                           (decay_matrix, UC_matrix, jac_indices, coop_ET_matrix, coop_jac_indices),
                           initial_pop, rtol=rtol, atol=atol, nsteps=nsteps, quiet=quiet)
