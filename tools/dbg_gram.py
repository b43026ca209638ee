import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from scedar_b200 import device as dev
from oracle import sdm_oracle as orc
rng=np.random.default_rng(0)
for n,p in ((8,4),(130,20),(500,20)):
    x=rng.random((n,p))
    c=dev.Cells.from_dense(torch.from_numpy(x).cuda(),"euclidean")
    d=c.pdist_symmetric().cpu().numpy()
    ref=orc.pdist(x,"euclidean")
    err=np.abs(d-ref)
    print(n,p,"max err",err.max(), "bad entries", (err>1e-9).sum())
    if err.max()>1e-9:
        bad=np.argwhere(err>1e-9)
        print(bad[:10].tolist())
        print("rows with bad:", np.unique(bad[:,0])[:40], "cols:", np.unique(bad[:,1])[:40])
