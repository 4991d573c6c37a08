import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import *
import wl_b200 as wl
def run(name, exact, nsteps=100, **kw):
    o, s = make_pair(exact=exact, **kw)
    for _ in range(nsteps): o.mom_step()
    wl.lib.check(s.flow.L, s.flow.L.wl_sim_step_n(s.flow.h, nsteps))
    dn = np.asarray(o.iters, int) - np.asarray(s.pois.n, int)
    print(name, "exact" if exact else "fmad", "u %.2e p %.2e" % (rel_l2(s.flow.u, o.field("u")), rel_l2(s.flow.p, o.field("p"))), "dn!=0:", int((dn != 0).sum()), "max|dn|", int(np.abs(dn).max()), "dt rel %.1e" % np.max(np.abs(o.dt - s.flow.Δt) / o.dt), "sum n", int(np.sum(o.iters)))
n = 32
u0 = tgv3d_u0((n + 2,) * 3, n); nu = float(F(1 / (2 * np.pi / n * 1600)))
for ex in (False, True):
    run("tgv32", ex, dims=(n,)*3, uBC=(0.,0.,0.), nu=nu, perdir=(1,2,3), u0=u0)
    run("sphere", ex, dims=(64,32,32), uBC=(1.,0.,0.), nu=8/100, sphere=((15.,15.,15.),4.), exitBC=True)
    run("circle", ex, dims=(96,64), uBC=(1.,0.), nu=16/100, sphere=((31.,31.),8.))
