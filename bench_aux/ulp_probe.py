import numpy as np, sys
sys.path.insert(0,".")
import covfn_b200 as cf
def ulps(got, want_ld):
    want = want_ld.astype(np.float64)
    return np.abs((got.astype(np.longdouble) - want_ld) / np.spacing(np.abs(want)).astype(np.longdouble)).astype(np.float64)
def entries(k, t):
    return cf.gramian(k, t.reshape(1, -1), np.zeros((1, 1))).Matrix()[:, 0]
rng = np.random.default_rng(0)
for hi in (2.0, 8.0, 20.0, 37.0):
    t = rng.uniform(0, hi, 40000); r2=(t*t).astype(np.longdouble)
    u = ulps(entries(cf.EQ(), t), np.exp(-r2/2)); print("EQ t<%g: max %.2f mean %.3f p99 %.2f"%(hi,u.max(),u.mean(),np.quantile(u,.99)))
t = rng.uniform(0, 300.0, 30000); r2=(t*t).astype(np.longdouble)
u = ulps(entries(cf.Exp(), t), np.exp(-np.sqrt(r2))); print("Exp: max %.2f mean %.3f"%(u.max(),u.mean()))
t = rng.uniform(0, 30.0, 30000); r2=(t*t).astype(np.longdouble)
u = ulps(entries(cf.Exp(), t), np.exp(-np.sqrt(r2))); print("Exp t<30: max %.2f mean %.3f"%(u.max(),u.mean()))
s=np.sqrt(5*r2); u=ulps(entries(cf.MaternP(2), t),(1+s+s*s/3)*np.exp(-s)); print("M2 t<30: max %.2f mean %.3f"%(u.max(),u.mean()))
t = rng.uniform(0, 1e3, 30000); r2=(t*t).astype(np.longdouble)
u=ulps(entries(cf.RQ(1), t), 1/(1+r2/2)); print("RQ1: max %.2f mean %.3f"%(u.max(),u.mean()))
u=ulps(entries(cf.RQ(3), t), (1+r2/6)**-3); print("RQ3: max %.2f mean %.3f"%(u.max(),u.mean()))
