"""simetuc_b200: B200-native microscopic rate-equation path behind simetuc's Python API.

    from simetuc_b200 import Simulations, configs
    sim = Simulations(configs.yb_tm_cte(), devices=[0])
    sol = sim.simulate_power_dependence([1e6, 1e8, 1e10])

Modules mirror the reference: `lattice`, `precalculate`, `odesolver`, `simulations`,
`optimize`, `settings`, `util`; `engine.RateSystem` owns the device-resident system and `_lib`
binds the C ABI (`include/simetuc_b200.h`).  Importing the package does not need a GPU; any
compute call does (there is no CPU fallback).
"""
VERSION = '0.1.0'
from .settings import Settings            # noqa: E402,F401
from .simulations import Simulations      # noqa: E402,F401
