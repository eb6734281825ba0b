import time, numpy as np, sys
sys.path.insert(0,".")
import covfn_b200 as cf
n,d=1<<17,8
rng=np.random.default_rng(0)
X=rng.standard_normal((n,d))/np.sqrt(d); y=rng.standard_normal(n)
res={}
for devs in ([0],[0,1],[0,1]):
    cf.init(devs)
    G=cf.gramian(cf.MaternP(2), X.T)
    A=1e-2*cf.I(n)+G
    x,it,r=A.solve(y,maxiter=10)
    true=np.linalg.norm(y-(A@x))
    print(devs, "recurrence %.10e true %.10e rel diff %.2e"%(r,true,abs(r-true)/true))
    res.setdefault(str(devs),[]).append(x)
print("2-device run-to-run identical:", np.array_equal(res["[0, 1]"][0],res["[0, 1]"][1]))
print("1 vs 2 device rel diff of x:", np.linalg.norm(res["[0]"][0]-res["[0, 1]"][0])/np.linalg.norm(res["[0]"][0]))
# one product: 1 vs 2 devices
cf.init([0]); b1=cf.gramian(cf.MaternP(2), X.T)@y
cf.init([0,1]); b2=cf.gramian(cf.MaternP(2), X.T)@y
print("product rel diff:", np.linalg.norm(b1-b2)/np.linalg.norm(b1))
