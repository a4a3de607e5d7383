import sys, numpy as np, torch
sys.path.insert(0, '.')
from ngacf_b200 import hostdata
from ngacf_b200.data import Interactions
from ngacf_b200.evaluate import AllNegEvaluator
dev='cuda:0'
U,I,E=29858,40981,1027370
u,i=hostdata.synth_bipartite(U,I,E,0); (tu,ti),(su,si)=hostdata.split_per_user(u,i,U,1)
inter=Interactions.from_arrays(U,I,tu,ti,su,si,device=dev)
Z=torch.randn(U+I,64,device=dev)*0.3
ev=AllNegEvaluator(inter,'tc')
for _ in range(3): ev.rank(Z)
torch.cuda.synchronize()
print('fallback', ev.n_fallback)
