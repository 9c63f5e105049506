"""Where a bench step spends its time: wall clock around every call of the device-resident step
versus the kernel time the library reports.  DEVELOPMENT TOOL."""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
import torch                                      # noqa: E402
from simetuc_b200 import lattice, precalculate   # noqa: E402
import bench                                      # noqa: E402

n = 32
cte = bench.make_cte()
powers = bench.powers_list()
lat = lattice.generate(cte, seed=2000)
system = precalculate.build_system(cte, lat, n_sys=n)
for k, P in enumerate(powers):
    for exc in cte['excitations']['NIR_980']:
        exc.power_dens = P
    system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=k)
N, kinds = system.N, system.sS + system.sA
_, _, pop = system.ground_states()
t_pulse, t_sol = bench.t_grids(cte)
f64 = dict(dtype=torch.float64, device='cuda')
d_tp = torch.tensor(t_pulse, **f64)
d_ts = torch.tensor(t_sol, **f64)
d_y0 = torch.empty(n, N, **f64)
system.to_planes(np.tile(pop.astype(np.float64), (n, 1)), d_y0.data_ptr())
d_yp = torch.empty(n, 2, N, **f64)
d_ys = torch.empty(n, 1000, N, **f64)
d_avg = torch.empty(n, 1000, kinds, **f64)
ref_of_plane = torch.tensor(bench._ref_of_plane(system), device='cuda')


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(4):
    for with_y in (True, False):
        t0 = sync()
        system.integrate_dev(d_tp.data_ptr(), 2, d_y0.data_ptr(), n, True, 1e-8, 1e-15, y_out_ptr=d_yp.data_ptr())
        t1 = sync()
        kp = system.last_timing()[0]
        y0r = d_yp[:, -1, :].index_select(1, ref_of_plane).contiguous()
        t2 = sync()
        system.integrate_dev(d_ts.data_ptr(), 1000, y0r.data_ptr(), n, False, 1e-8, 1e-15,
                             y_out_ptr=d_ys.data_ptr() if with_y else None, avg_out_ptr=d_avg.data_ptr())
        t3 = sync()
        kr = system.last_timing()[0]
        print('rep %d with_y=%d: pulse wall %.2f ms (kernel %.2f) | select %.2f | relax wall %.2f ms (kernel %.2f)'
              % (rep, with_y, 1e3 * (t1 - t0), kp, 1e3 * (t2 - t1), 1e3 * (t3 - t2), kr))

# ---- the bench's own loop, with and without the L2 flush ------------------------------------
ext = torch.cuda.ExternalStream(system.stream, device=torch.device('cuda', 0))
flush = torch.empty(256 * 1024 * 1024 // 8, **f64)
for do_flush in (0, 1, 2, 0, 1, 2):
    sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t0 = time.perf_counter()
    ev[0].record(ext)
    if do_flush == 1:
        with torch.cuda.stream(ext):
            flush.zero_()
    elif do_flush == 2:
        flush.zero_()          # torch's current stream, not the library's
    ev[1].record(ext)
    system.integrate_dev(d_tp.data_ptr(), 2, d_y0.data_ptr(), n, True, 1e-8, 1e-15, y_out_ptr=d_yp.data_ptr())
    kp = system.last_timing()[0]
    with torch.cuda.stream(ext):
        y0r = d_yp[:, -1, :].index_select(1, ref_of_plane).contiguous()
    ev[2].record(ext)
    system.integrate_dev(d_ts.data_ptr(), 1000, y0r.data_ptr(), n, False, 1e-8, 1e-15,
                         y_out_ptr=d_ys.data_ptr(), avg_out_ptr=d_avg.data_ptr())
    kr = system.last_timing()[0]
    ev[3].record(ext)
    t1 = sync()
    print('flush=%d: wall %.1f ms; events: flush %.2f, pulse %.2f (kernel %.2f), relax %.2f (kernel %.2f)' % (
        do_flush, 1e3 * (t1 - t0), ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), kp,
        ev[2].elapsed_time(ev[3]), kr))
