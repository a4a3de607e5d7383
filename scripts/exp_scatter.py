import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from ngacf_b200 import ops
dev='cuda:0'
U,I,B=29858,40981,2048
Z=torch.randn(U+I,64,device=dev)
G=torch.zeros_like(Z)
def bench(users, items, tag):
    d=torch.randn(B,device=dev)
    for _ in range(3): ops.score_pairs_bwd(Z,U,users,items,d,G)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.score_pairs_bwd(Z,U,users,items,d,G)
    e1.record(); torch.cuda.synchronize()
    print(tag, e0.elapsed_time(e1)/20*1000,'us')
rand_u=torch.randint(0,U,(B,),device=dev); rand_i=torch.randint(0,I,(B,),device=dev)
bench(rand_u,rand_i,'random users/items')
sorted_u=torch.sort(torch.randint(0,75,(B,),device=dev))[0]
bench(sorted_u,rand_i,'75 sorted users, random items')
bench(sorted_u,torch.randint(0,50,(B,),device=dev),'75 users, 50 items')
bench(torch.zeros(B,dtype=torch.int64,device=dev),rand_i,'one user')
