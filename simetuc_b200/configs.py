"""Parameter sets (data, not code) of the reference's shipped configurations, as `cte`
dictionaries.  The numbers are the physical constants of the reference's example files:

* `test_cte()`      - `simetuc/test/conftest.py:26-117` (7 Tm states, Vis_473 active)
* `sim_cte()`       - `conftest.py:288-312` (the set that generated `y_sol_2S_2A.hdf5`)
* `yb_tm_cte()`     - `simetuc/config_file_Yb_Tm.cfg:10-248` (9 Tm states, NIR_980 active)
* `esa_cte()`       - `simetuc/config_file_ESA.cfg` shape: 3-level ESA, no energy transfer

Every builder takes `types=` (a namespace providing `IonType, Transition, DecayTransition,
Excitation, EneryTransferProcess`); the default is this package's `util`.  Tests pass the
reference's own `simetuc.util` to feed the identical numbers to the unmodified reference.
"""
from collections import OrderedDict

from . import util as _util
from .settings import Settings

BNAYF4 = {'name': 'bNaYF4', 'a': 5.9738, 'b': 5.9738, 'c': 3.5297,
          'alpha': 90.0, 'beta': 90.0, 'gamma': 120.0, 'spacegroup': 'P-6',
          'sites_pos': [(0.0, 0.0, 0.0), (2 / 3, 1 / 3, 0.5)], 'sites_occ': [1.0, 0.5],
          'cell_par': [5.9738, 5.9738, 3.5297, 90.0, 90.0, 120.0]}


def _lattice(N_uc, S_conc, A_conc, **extra):
    lat = dict(BNAYF4)
    lat.update({'N_uc': N_uc, 'S_conc': S_conc, 'A_conc': A_conc})
    lat.update(extra)
    return lat


def test_cte(types=_util):
    """`conftest.py:26-117`."""
    T, D, E, P, I = (types.Transition, types.DecayTransition, types.Excitation,
                     types.EneryTransferProcess, types.IonType)
    S, A = I.S, I.A
    et = OrderedDict()
    et['CR50'] = P([T(A, 5, 3), T(A, 0, 2)], mult=6, strength=887920884.0, name='CR50')
    et['ETU53'] = P([T(A, 5, 6), T(A, 3, 1)], mult=6, strength=450220614.0, name='ETU53')
    et['ETU55'] = P([T(A, 5, 6), T(A, 5, 4)], mult=6, strength=0.0, name='ETU55')
    et['BackET'] = P([T(A, 3, 0), T(S, 0, 1)], mult=6, strength=4502.20614, name='BackET')
    et['EM'] = P([T(S, 1, 0), T(S, 0, 1)], mult=6, strength=45022061400.0, name='EM')
    et['ETU1'] = P([T(S, 1, 0), T(A, 0, 2)], mult=6, strength=10000.0, name='ETU1')
    et['coop1'] = P([T(S, 1, 0), T(S, 1, 0), T(A, 0, 5)], mult=6, strength=1000.0, name='coop1')
    decay = {
        'branching_A': {D(A, 2, 1, branching_ratio=0.4), D(A, 3, 1, branching_ratio=0.3),
                        D(A, 4, 3, branching_ratio=0.999), D(A, 5, 1, branching_ratio=0.15),
                        D(A, 5, 2, branching_ratio=0.16), D(A, 5, 3, branching_ratio=0.04),
                        D(A, 5, 4, branching_ratio=0.0), D(A, 6, 1, branching_ratio=0.43)},
        'branching_S': {D(S, 1, 0, branching_ratio=1.0)},
        'decay_A': {D(A, 1, 0, decay_rate=83.33333333333333), D(A, 2, 0, decay_rate=40000.0),
                    D(A, 3, 0, decay_rate=500.0), D(A, 4, 0, decay_rate=500000.0),
                    D(A, 5, 0, decay_rate=1315.7894736842104),
                    D(A, 6, 0, decay_rate=14814.814814814814)},
        'decay_S': {D(S, 1, 0, decay_rate=400.0)}}
    exc = OrderedDict()
    exc['NIR_1470'] = [E(A, 5, 6, False, 9 / 5, 2e-4, 1e7, 1e-8)]
    exc['NIR_800'] = [E(A, 0, 3, False, 13 / 9, 0.0044, 1e7, 1e-8),
                      E(A, 2, 5, False, 11 / 9, 0.002, 1e7, 1e-8)]
    exc['NIR_980'] = [E(S, 0, 1, False, 4 / 3, 0.0044, 1e7, 1e-8)]
    exc['Vis_473'] = [E(A, 0, 5, True, 13 / 9, 0.00093, 1e6, 1e-8)]
    return Settings({
        'version': 1, 'energy_transfer': et, 'decay': decay, 'excitations': exc,
        'lattice': _lattice(20, 0.3, 0.3, d_max=100.0, d_max_coop=25.0),
        'no_console': True, 'no_plot': True,
        'simulation_params': {'N_steps': 1000, 'N_steps_pulse': 2, 'atol': 1e-15, 'rtol': 1e-3},
        'states': {'activator_ion_label': 'Tm', 'activator_states': 7,
                   'activator_states_labels': ['3H6', '3F4', '3H5', '3H4', '3F3', '1G4', '1D2'],
                   'sensitizer_ion_label': 'Yb', 'sensitizer_states': 2,
                   'sensitizer_states_labels': ['GS', 'ES']}})


def sim_cte(types=_util):
    """`conftest.py:288-312`: `test_cte` with one more branching ratio and 5 pair processes."""
    T, D, P, I = types.Transition, types.DecayTransition, types.EneryTransferProcess, types.IonType
    S, A = I.S, I.A
    cte = test_cte(types)
    cte['decay']['branching_A'].add(D(A, 3, 2, branching_ratio=0.1))
    et = OrderedDict()
    et['CR50'] = P([T(A, 5, 3), T(A, 0, 2)], mult=6, strength=2893199540.0, name='CR50')
    et['ETU53'] = P([T(A, 5, 6), T(A, 3, 1)], mult=6, strength=254295690.0, name='ETU53')
    et['BackET'] = P([T(A, 3, 0), T(S, 0, 1)], mult=6, strength=4502.20614, name='BackET')
    et['EM'] = P([T(S, 1, 0), T(S, 0, 1)], mult=6, strength=45022061400.0, name='EM')
    et['ETU1'] = P([T(S, 1, 0), T(A, 0, 2)], mult=6, strength=10000.0, name='ETU1')
    cte['energy_transfer'] = et
    return cte


def yb_tm_cte(N_uc=15, S_conc=5.0, A_conc=0.5, power_dens=1e7, t_pulse=1e-8,
              activator_states=9, rtol=1e-8, atol=1e-15, N_steps=1000, types=_util):
    """`config_file_Yb_Tm.cfg`: beta-NaYF4:Yb,Tm under 980 nm excitation of Yb.

    `t_pulse=None` makes the power dependence a CW steady state (`simulations.py:1188-1193`).
    `activator_states=7` drops 1I6/3P0 and the processes that need them (the reference
    discards transitions naming non-existent states, `precalculate.py:142-147,172-176`)."""
    T, D, E, P, I = (types.Transition, types.DecayTransition, types.Excitation,
                     types.EneryTransferProcess, types.IonType)
    S, A = I.S, I.A
    nA = activator_states
    et = OrderedDict()

    def add(name, trans, strength, mult=6):
        if all(max(t.state_i, t.state_f) < (nA if t.ion is A else 2) for t in trans):
            et[name] = P(trans, mult=mult, strength=strength, name=name)

    add('CR50', [T(A, 5, 3), T(A, 0, 2)], 3.035e+09)
    add('CR30', [T(A, 3, 1), T(A, 0, 1)], 1.318e+08)
    add('ETU53', [T(A, 5, 6), T(A, 3, 1)], 1.821e+11)
    add('ETU55_1I6', [T(A, 5, 2), T(A, 5, 7)], 1.603e+06)
    add('BackET', [T(A, 3, 0), T(S, 0, 1)], 1.312e+07)
    add('EM', [T(S, 1, 0), T(S, 0, 1)], 4.199e+07)
    add('ETU1', [T(S, 1, 0), T(A, 0, 2)], 7.377e+07)
    add('ETU2', [T(S, 1, 0), T(A, 1, 4)], 1.563e+08)
    add('ETU3', [T(S, 1, 0), T(A, 3, 5)], 1.142e+10)
    add('ETU4', [T(S, 1, 0), T(A, 5, 6)], 0.0)
    add('coop1', [T(S, 1, 0), T(S, 1, 0), T(A, 0, 5)], 0.0, mult=8)
    add('coop2', [T(S, 0, 1), T(S, 0, 1), T(A, 5, 0)], 0.0, mult=8)
    lifetimes = [12e-3, 25e-6, 2e-3, 2e-6, 775e-6, 67.5e-6, 101.8e-6, 5e-6]
    branching = [(2, 1, 0.4), (3, 1, 0.3), (3, 2, 0.1), (4, 3, 0.999), (5, 1, 0.15),
                 (5, 2, 0.16), (5, 3, 0.04), (5, 4, 0.02), (6, 1, 0.43), (7, 1, 0.6),
                 (7, 3, 0.16), (7, 5, 0.14), (7, 6, 0.0), (8, 7, 1.0)]
    decay = {
        'decay_S': {D(S, 1, 0, decay_rate=1 / 2.5e-3)},
        'decay_A': {D(A, i + 1, 0, decay_rate=1 / tau) for i, tau in enumerate(lifetimes)
                    if i + 1 < nA},
        'branching_S': set(),
        'branching_A': {D(A, i, f, branching_ratio=b) for i, f, b in branching if i < nA}}
    exc = OrderedDict()
    exc['Vis_473'] = [E(A, 0, 5, False, 13 / 9, 9.3e-3, 1e6, 5e-9)]
    exc['NIR_1470'] = [E(A, 5, 6, False, 9 / 5, 2e-4, 1e6, 1e-8)]
    exc['NIR_980'] = [E(S, 0, 1, True, 4 / 3, 4.4e-3, power_dens, t_pulse)]
    exc['NIR_800'] = [E(A, 0, 3, False, 13 / 9, 4.4e-3, 1e2, 1e-8),
                      E(A, 2, 5, False, 11 / 9, 4e-3, 1e2, 1e-8)]
    labels = ['3H6', '3F4', '3H5', '3H4', '3F3', '1G4', '1D2', '1I6', '3P0'][:nA]
    return Settings({
        'version': 1, 'energy_transfer': et, 'decay': decay, 'excitations': exc,
        'lattice': _lattice(N_uc, S_conc, A_conc, d_max=100.0, d_max_coop=50.0),
        'no_console': True, 'no_plot': True,
        'simulation_params': {'N_steps': N_steps, 'N_steps_pulse': 2, 'atol': atol, 'rtol': rtol},
        'power_dependence': [1e6, 1e13, 8],
        'optimization': {'processes': [et[k] for k in ('ETU1', 'ETU2', 'ETU3', 'BackET', 'EM')
                                       if k in et],
                         'method': 'COBYLA'},
        'states': {'activator_ion_label': 'Tm', 'activator_states': nA,
                   'activator_states_labels': labels,
                   'sensitizer_ion_label': 'Yb', 'sensitizer_states': 2,
                   'sensitizer_states_labels': ['GS', 'ES']}})


def esa_cte(N_uc=50, A_conc=0.3, power_dens=1e6, t_pulse=None, rtol=1e-8, atol=1e-15,
            types=_util):
    """Three-level ESA system without energy transfer (shape of `config_file_ESA.cfg` and of
    `test_simulations.py:347-407`): GS -> ES1 -> ES2 pumped by one source, decay rates k1, k2."""
    T, D, E, I = types.Transition, types.DecayTransition, types.Excitation, types.IonType
    A = I.A
    decay = {'decay_S': set(), 'branching_S': set(),
             'decay_A': {D(A, 1, 0, decay_rate=1e3), D(A, 2, 0, decay_rate=1e6)},
             'branching_A': {D(A, 2, 1, branching_ratio=0.0)}}
    exc = OrderedDict()
    exc['ESA'] = [E(A, 0, 1, True, 0.0, 1e-3, power_dens, t_pulse),
                  E(A, 1, 2, True, 0.0, 1e-3, power_dens, t_pulse)]
    return Settings({
        'version': 1, 'energy_transfer': OrderedDict(), 'decay': decay, 'excitations': exc,
        'lattice': _lattice(N_uc, 0.0, A_conc),
        'no_console': True, 'no_plot': True,
        'simulation_params': {'N_steps': 1000, 'N_steps_pulse': 2, 'atol': atol, 'rtol': rtol},
        'states': {'activator_ion_label': 'Tm', 'activator_states': 3,
                   'activator_states_labels': ['GS', 'ES1', 'ES2'],
                   'sensitizer_ion_label': 'Yb', 'sensitizer_states': 0,
                   'sensitizer_states_labels': []}})
