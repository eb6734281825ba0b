import time, numpy as np, sys, torch
sys.path.insert(0,".")
import covfn_b200 as cf
n=1<<20
rng=np.random.Generator(np.random.Philox(1))
X=rng.standard_normal((n,3)); a=rng.standard_normal(n)
a_pin=torch.from_numpy(a).pin_memory(); b_pin=torch.empty(n,dtype=torch.float64).pin_memory()
a_np=a_pin.numpy(); b_np=b_pin.numpy()
XT=X.T
k=cf.EQ()
for it in range(4):
    t0=time.perf_counter(); G=cf.gramian(k, XT).set_row_range(0,n); t1=time.perf_counter(); G.handle(); t2=time.perf_counter()
    cf.mul_(b_np,G,a_np); t3=time.perf_counter(); ms,_=G.last_timing(); G.close(); t4=time.perf_counter()
    print("ctor %.1f create %.1f mul %.1f (kernel %.1f) close %.1f total %.1f"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,ms,(t4-t3)*1e3,(t4-t0)*1e3))
# pageable
b2=np.zeros(n)
for it in range(2):
    t0=time.perf_counter(); G=cf.gramian(k, XT); G.handle(); t2=time.perf_counter(); cf.mul_(b2,G,a); t3=time.perf_counter(); ms,_=G.last_timing(); G.close(); t4=time.perf_counter()
    print("pageable: create %.1f mul %.1f (kernel %.1f) close %.1f"%((t2-t0)*1e3,(t3-t2)*1e3,ms,(t4-t3)*1e3))
