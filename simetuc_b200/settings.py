"""`cte`: the settings dictionary every layer of the rate-equation path reads.

The reference parses a YAML `.cfg` with the external `settings_parser` package
(`settings.py:405-485`) into a dict-like `Settings` with attribute access.  Text parsing is
outside the hot path; this module provides the same container
(`Settings`, `Settings.load_from_dict`) plus the package defaults of
`simetuc/config/settings.cfg:1-5` so that a `cte` built for the reference drops in unchanged.
"""
import copy
from typing import Dict, List

import numpy as np

#: defaults of the reference (`config/settings.cfg:1-5`)
DEFAULT_SIMULATION_PARAMS = {'rtol': 1e-5, 'atol': 1e-20, 'N_steps_pulse': 2, 'N_steps': 1000}


class Settings(dict):
    """Dictionary with attribute access (`cte.lattice['N_uc']`, `cte['no_console']`)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as err:
            raise AttributeError(name) from err

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]

    def __deepcopy__(self, memo):
        return Settings({k: copy.deepcopy(v, memo) for k, v in self.items()})

    @classmethod
    def load_from_dict(cls, d: Dict) -> 'Settings':
        """Wrap a plain dict, filling solver defaults (reference: `settings.py:334-359`)."""
        cte = cls(d)
        params = dict(DEFAULT_SIMULATION_PARAMS)
        params.update(cte.get('simulation_params') or {})
        cte['simulation_params'] = params
        cte.setdefault('no_console', True)
        cte.setdefault('no_plot', True)
        return cte


def parse_power_dependence(user_list) -> List[float]:
    """`[min, max, n]` -> n log-spaced power densities (reference: `settings.py:362-369`)."""
    lo, hi, n = user_list[0], user_list[1], int(user_list[2])
    return list(np.logspace(np.log10(lo), np.log10(hi), n))


def parse_conc_dependence(conc_lists, N_uc, N_uc_list=None) -> Dict:
    """`[[S...], [A...]]` -> every (S, A) combination, S varying fastest (the reference's
    `np.meshgrid(S, A)` flattened row-major, `settings.py:372-402`); a short `N_uc_list` is
    padded with its last element."""
    S_list, A_list = conc_lists
    concs = [(float(s), float(a)) for a in A_list for s in S_list]
    N_uc_list = [N_uc] * len(concs) if N_uc_list is None else list(N_uc_list)
    if len(N_uc_list) > len(concs):
        raise ValueError('N_uc_list has more elements than concentrations.')
    N_uc_list += [N_uc_list[-1]] * (len(concs) - len(N_uc_list))
    return {'concentrations': concs, 'N_uc_list': N_uc_list}
