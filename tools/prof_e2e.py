"""Timing of the end-to-end API call the bench times, in a process that has torch loaded like the
bench has (DEVELOPMENT TOOL)."""
import copy
import gc
import sys
import time
sys.path.insert(0, '.')
import torch                                      # noqa: E402,F401
import bench                                      # noqa: E402
from simetuc_b200 import engine                   # noqa: E402
from simetuc_b200.simulations import Simulations  # noqa: E402

acc = {'gc': 0.0, 'deepcopy': 0.0, 'integrate': 0.0, 'n_gc': 0}
_t = [0.0]


def gc_cb(phase, info):
    if phase == 'start':
        _t[0] = time.perf_counter()
    else:
        acc['gc'] += time.perf_counter() - _t[0]
        acc['n_gc'] += 1


gc.callbacks.append(gc_cb)
_deepcopy = copy.deepcopy
depth = [0]


def timed_deepcopy(x, memo=None, _nil=[]):
    if depth[0]:
        return _deepcopy(x, memo)
    depth[0] = 1
    t0 = time.perf_counter()
    try:
        return _deepcopy(x, memo)
    finally:
        acc['deepcopy'] += time.perf_counter() - t0
        depth[0] = 0


copy.deepcopy = timed_deepcopy
_integrate = engine.RateSystem.integrate


def timed_integrate(self, *a, **k):
    t0 = time.perf_counter()
    try:
        return _integrate(self, *a, **k)
    finally:
        acc['integrate'] += time.perf_counter() - t0


engine.RateSystem.integrate = timed_integrate
powers = bench.powers_list()
sim = Simulations(bench.make_cte(), devices=[0])
sim.cte['lattice_seed'] = 2000
for rep in range(6):
    for k in acc:
        acc[k] = 0
    t0 = time.perf_counter()
    sol = sim.simulate_power_dependence(powers, keep_y=False)
    _ = sol.populations
    print('e2e call: %.1f ms' % (1e3 * (time.perf_counter() - t0)),
          {k: (round(1e3 * v, 1) if k != 'n_gc' else v) for k, v in acc.items()})
