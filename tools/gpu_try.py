import sys, time, glob, os
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
torch.set_default_dtype(torch.float64)
import helpers
names = sys.argv[1:] or ['hk_as5_chi002','hk_as5_chi000','hk_1d','hk_as5_rot','hk_methylium','hk_as24_rot','hk_as60','hk_as60_rot']
dev = torch.device('cuda:0')
for name in names:
    g = helpers.load_golden(name)
    pot = helpers.potential_from_golden(g)
    pr = helpers.propagator_from_golden(g, dev)
    nt, dt, e0 = int(g['nt']), float(g['dt']), float(g['energy0_es'])
    # step-by-step (drop-in API)
    auto = np.zeros(nt, complex); ic = np.zeros(nt, complex)
    t0 = time.time()
    for k in range(nt):
        auto[k] = pr.autocorrelation(e0); ic[k] = pr.ic_correlation(pot, e0); pr.step(pot, dt)
    torch.cuda.synchronize(); t1 = time.time()
    y = pr.y.cpu().numpy(); c = pr.c.cpu().numpy(); sg = pr.sign_trackers['prefactorC']['signs'].real.cpu().numpy()
    nk = g['y_final'].shape[1]
    print(f"{name:16s} [{pr.kernel_name()}] auto {helpers.relerr(auto,g['autocorrelation']):.2e} ic {helpers.relerr(ic,g['ic_correlation']):.2e} "
          f"y {helpers.relerr(y[:,:nk],g['y_final']):.2e} c {helpers.relerr(c,g['c_final']):.2e} signs {(sg==g['signs_C']).all()}  {t1-t0:.2f}s")
    # fused
    pr2 = helpers.propagator_from_golden(g, dev)
    a0 = pr2.autocorrelation(e0); i0 = pr2.ic_correlation(pot, e0)
    t0 = time.time()
    a, i = pr2.propagate(pot, dt, nt-1, e0)
    torch.cuda.synchronize(); t1 = time.time()
    a = np.concatenate(([a0], a)); i = np.concatenate(([i0], i))
    print(f"{'':16s} fused: auto {helpers.relerr(a,g['autocorrelation']):.2e} ic {helpers.relerr(i,g['ic_correlation']):.2e}  {t1-t0:.3f}s")
