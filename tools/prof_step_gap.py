"""Where the device-resident bench step spends the time that is not inside the two integrator
kernels (DEVELOPMENT TOOL)."""
import sys
import time
sys.path.insert(0, '.')
import torch                 # noqa: E402
import bench                 # noqa: E402

dev = torch.device('cuda', 0)
cte, lat, system = bench.build_sweep_system(0, list(range(32)))
t_pulse, t_sol = bench.t_grids(cte)
sweep = bench.DeviceSweep(system, 32, dev, t_pulse, t_sol)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
for _ in range(3):
    sweep.step()
torch.cuda.synchronize()
for label, do_flush, sampler in (('plain', False, False), ('flush', True, False), ('flush+sampler', True, True)):
    smp = bench.ClockSampler(0) if sampler else None
    if smp:
        smp.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks = []
    t0 = time.perf_counter()
    ev0.record(sweep.ext)
    for _ in range(10):
        if do_flush:
            with torch.cuda.stream(sweep.ext):
                flush.zero_()
        ms_p, ms_r, *_ = sweep.step()
        ks.append(ms_p + ms_r)
    ev1.record(sweep.ext)
    torch.cuda.synchronize()
    wall = 1e2 * (time.perf_counter() - t0)
    if smp:
        smp.stop()
    print('%-14s event %.2f ms/step, wall %.2f ms/step, kernels %.2f ms/step' % (
        label, ev0.elapsed_time(ev1) / 10, wall, sum(ks) / 10), flush=True)
import os
os.environ['STB_TIMING'] = '1'
sweep.step()
