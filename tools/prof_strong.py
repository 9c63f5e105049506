"""Strong-scaling shards of the C2 sweep on one GPU: 32 / 16 / 8 / 4 systems with the cluster sizes
the library picks and with forced ones (DEVELOPMENT TOOL)."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
import bench                                     # noqa: E402

for n_gpus in (1, 2, 4, 8):
    mine = list(range(0, 32, n_gpus))            # rank 0's shard (holds the slowest system? no: the fastest)
    mine = list(range(n_gpus - 1, 32, n_gpus))   # last rank's shard: holds power #31, the slowest system
    cte, lat, system = bench.build_sweep_system(0, mine)
    _, _, pop = system.ground_states()
    t_pulse, t_sol = bench.t_grids(cte)
    y0 = np.tile(pop.astype(float), (len(mine), 1))
    for forced in (None, 4, 8):
        os.environ.pop('STB_GROUP_CTAS', None)
        if forced:
            os.environ['STB_GROUP_CTAS'] = str(forced)
        best = 1e9
        for rep in range(2):
            _, _, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol, y0, rtol=1e-8, atol=1e-15, want_y=False)
            best = min(best, system.last_timing()[0])
        prof = system.last_profile()
        print('N=%d (%2d systems): forced G=%s -> G=%d clusters=%d  %.2f ms  (%.0f solves/s for the sweep)' % (
            n_gpus, len(mine), forced, prof['G'], prof['groups'], best, 32e3 / best), flush=True)
    system.close()
