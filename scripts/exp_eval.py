import sys, numpy as np, torch
sys.path.insert(0, '.')
from ngacf_b200 import ops, hostdata, _lib
from ngacf_b200.data import Interactions
from ngacf_b200.evaluate import AllNegEvaluator
dev='cuda:0'
U,I,E=29858,40981,1027370
u,i=hostdata.synth_bipartite(U,I,E,0); (tu,ti),(su,si)=hostdata.split_per_user(u,i,U,1)
inter=Interactions.from_arrays(U,I,tu,ti,su,si,device=dev)
Z=torch.randn(U+I,64,device=dev)*0.3
for mode in ('tc','exact'):
    ev=AllNegEvaluator(inter,mode)
    ev.rank(Z); torch.cuda.synchronize()
    _lib.PROFILE=[]
    for _ in range(3): ev.rank(Z); ev.metrics()
    torch.cuda.synchronize()
    agg={}
    for name,args,e0,e1 in _lib.PROFILE: agg.setdefault(name,[]).append(e0.elapsed_time(e1))
    _lib.PROFILE=None
    print(mode, {k: round(float(np.mean(v))*1000,1) for k,v in agg.items()}, 'us; fallback', ev.n_fallback)
