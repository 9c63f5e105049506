"""Cell-list pair build against the O(n^2) scan on a large cutoff lattice (DEVELOPMENT TOOL)."""
import os
import sys
import time
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice        # noqa: E402
from simetuc_b200.engine import RateSystem, pair_processes   # noqa: E402

for n_uc, S in ((40, 25.0), (60, 50.0)):
    cte = configs.yb_tm_cte(N_uc=n_uc, S_conc=S, A_conc=0.3)
    lat = lattice.generate(cte, seed=4000)
    for cells in ('1', '0'):
        os.environ['STB_CELLS'] = cells
        for rep in range(2):
            t0 = time.perf_counter()
            system = RateSystem(lat, dict(pair_processes(cte['energy_transfer'])), d_max=12.0, geometry='positions')
            dt = time.perf_counter() - t0
            info = (system.n_ions, system.n_pairs, system.get_option('cell_lists'))
            system.close()
        print('N_uc %d %g%% Yb: %d ions, %d pairs within 12 A, cell lists %d: build %.1f ms' % (
            n_uc, S, info[0], info[1], info[2], 1e3 * dt), flush=True)
