import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import svbrdf_estimation_b200 as S
from oracle import reference_port as O
from tests.common import synthetic_maps
dev=torch.device('cuda',0)
for batch,size,nr,ns in ((4,128,3,6),(8,256,3,6),(4,256,9,18)):
    inp=synthetic_maps(batch,size,11 if size==128 else 21)
    torch.manual_seed(313); cfg=O.sample_loss_configs(batch,nr,ns)
    with torch.no_grad():
        r64=O.render_batch(inp.double().to(dev),cfg).cpu().numpy(); r32=O.render_batch(inp.to(dev),cfg).cpu().numpy()
    r=S.render_records(inp.to(dev),cfg).cpu().numpy()
    def stats(r):
        rel=np.abs(r-r64)/np.maximum(np.abs(r64),1e-30)
        dl=np.abs(np.log(r.astype(np.float64)+0.1)-np.log(r64+0.1))
        return "relL2 %.2e p99.9 %.1e p99.99 %.1e maxdlog %.1e mean %.2e"%(np.linalg.norm((r-r64).ravel())/np.linalg.norm(r64.ravel()),np.quantile(rel,0.999),np.quantile(rel,0.9999),dl.max(),rel.mean())
    print("B%d %d N%d  ours: %s | ref32: %s"%(batch,size,nr+ns,stats(r),stats(r32)))
