#!/usr/bin/env python
"""One-GPU check of the shape the 8-GPU doc-sharded bench runs on every rank: shard 3 of 8 of the C2 corpus, the
replicated 8192-query batch, shard lists cut to m = 282 — which must be the first 282 entries of the top-1000 lists."""
import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import bench, mse_b200
from mse_b200 import _native, synthetic
dev=torch.device('cuda',0)
c=bench.gen_corpus(dev)
to,pd,pt,dl,base=bench.shard_corpus(c,3,8)
nat=_native.NativeIndex(0); nat.bm25_load(to,pd,pt,dl,c.idf,c.avgdl,doc_base=base)
q=[torch.from_numpy(a).to(dev) for a in synthetic.make_bm25_queries(c,8192,seed=99)]
d1,s1,c1=nat.bm25_search(q[0],q[1],q[2],282,0.0)
d2,s2,c2=nat.bm25_search(q[0],q[1],q[2],1000,0.0)
torch.cuda.synchronize()
ok=bool(torch.equal(torch.clamp(c2,max=282),c1)); v=torch.arange(282,device=dev)[None,:]<c1[:,None]
print('n8 shard check: counts',ok,'docs',bool(torch.equal(d1[v],d2[:,:282][v])),'scores',bool(torch.equal(s1[v],s2[:,:282][v])), nat.bm25_stats())
