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

import os as _os

# Every handle of the library owns a CUDA stream, and sweeps keep several handles in flight per
# device.  With the default of 8 hardware work queues, streams share queues and a short kernel of
# one handle (pair build, layout conversion) waits behind the persistent integrator kernel of
# another one; 32 is the maximum.  Only effective before the CUDA context of the process exists.
_os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
