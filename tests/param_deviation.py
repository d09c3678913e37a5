import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle
from dnbc_scala_b200 import _native as nat
from tests.util import *
def rel(a,b,floor):
    a=np.asarray(a); b=np.asarray(b)
    m=np.abs(b)>floor
    return float(np.max(np.abs(a[m]-b[m])/np.abs(b[m]))) if m.any() else 0.0
for (N,V,K,S,T) in [(10,[10]*5,[3,1,4,2,3],64,200),(12,[4],[],48,200),(10,[],[3]*5,64,200),(64,[10]*4,[8]*4,24,200),(64,[10]*2,[8]*4,96,1000),(64,[],[8,3,5],256,500)]:
    m = random_model(oracle, N, V, K, 61, unit_var=False)
    data = sample(oracle, m, [T]*S, 62)
    q = perturb(oracle, m, 63)
    want = q.copy()
    ll_want = oracle.em_iterate(want, data, 4)
    with ctx_for(nat, q) as ctx:
        upload(ctx, data, labels=False, dtype=np.float32)
        ll = ctx.em_iterate(4)
        p = ctx.get_params()
    out={"N":N,"ll":float(np.max(np.abs(ll-ll_want)/np.abs(ll_want))),
         "pi":rel(p["pi"],want.pi,1e-3),"A":rel(p["A"],want.A,1e-3),"B":rel(p["B"],want.B,1e-3) if V else 0}
    for c in range(len(K)):
        k=K[c]
        out.setdefault("w",0); out["w"]=max(out["w"],rel(p["w"][c,:,:k],want.w[c,:,:k],1e-3))
        out.setdefault("mu",0); out["mu"]=max(out["mu"],float(np.max(np.abs(p["mu"][c,:,:k]-want.mu[c,:,:k])/np.maximum(np.abs(want.mu[c,:,:k]),1.0))))
        out.setdefault("var",0); out["var"]=max(out["var"],rel(p["var"][c,:,:k],want.var[c,:,:k],1e-3))
    print({k:(v if k=="N" else float(f"{v:.2e}")) for k,v in out.items()})
