"""One process driving two GPUs through the public API (`Simulations(devices=[0, 1])`): results must
equal the single-device ones.  Run on a box with >= 2 GPUs (gpurun --gpus 2).  MEASUREMENT / CHECK TOOL."""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
import bench                                                   # noqa: E402
from simetuc_b200 import _lib, configs, precalculate            # noqa: E402
from simetuc_b200.simulations import Simulations               # noqa: E402

n_dev = _lib.load().stb_device_count()
print('devices:', n_dev)
assert n_dev >= 2
powers = bench.powers_list()
res = {}
for devs in ([0], [0, 1]):
    sim = Simulations(bench.make_cte(), devices=devs)
    sim.cte['lattice_seed'] = 2000
    sim.simulate_power_dependence(powers, keep_y=False)
    t0 = time.perf_counter()
    sol = sim.simulate_power_dependence(powers, keep_y=False)
    dt = time.perf_counter() - t0
    res[len(devs)] = sol.populations
    print('power dependence on %s: %.1f ms -> %.0f solves/s' % (devs, 1e3 * dt, 32 / dt))
assert np.array_equal(res[1], res[2]), np.abs(res[1] - res[2]).max()
concs = [(10.0, 0.5), (25.0, 0.3), (25.0, 1.0), (50.0, 2.0)] * 4
cte = configs.yb_tm_cte(N_uc=10, S_conc=25.0, A_conc=0.3, power_dens=1e7, rtol=1e-8, atol=1e-15, N_steps=300)
cte['lattice_seed'] = 77
out = {}
for devs in ([0], [0, 1]):
    precalculate.clear_system_pool()
    sim = Simulations(cte, devices=devs, streams_per_device=8)
    t0 = time.perf_counter()
    sol = sim.simulate_concentration_dependence(concs, dynamics=True, keep_y=False)
    dt = time.perf_counter() - t0
    out[len(devs)] = np.array([np.array(s.list_avg_data)[:, -1] for s in sol])
    print('concentration dependence (16 lattices) on %s: %.2f s' % (devs, dt))
assert np.array_equal(out[1], out[2])
sim = Simulations(cte, devices=[0, 1])
avg = sim.sample_simulation(sim.simulate_dynamics, N_samples=4)
print('sample_simulation over 2 devices ok:', len(avg.list_avg_data_ofs))
print('MULTI-DEVICE OK')
