import sys, numpy as np
sys.path.insert(0,'.')
import nerf_or_nothing_b200 as nb
from oracle import oracle as orc
from tests.gpu_util import configs_pair, dev, rel_err
def run(precision, R, kw):
    ncfg, ocfg = configs_pair(n_rays=R, precision=nb.PRECISIONS[precision], **kw)
    m = nb.AcceleratedMipNeRF(ncfg)
    S=ncfg.n_samples; M=R*S
    rng=np.random.default_rng(4)
    P,Dd=6*ncfg.deg_point,3+6*ncfg.deg_view
    params=orc.init_params(ocfg,7)
    nb_=sum(orc.layer_shapes(ocfg)[0]); params[-nb_:]=rng.normal(size=nb_).astype(np.float32)*0.1
    m.set_params(params)
    ep=rng.uniform(-1,1,(M,P)).astype(np.float32); ed=rng.uniform(-1,1,(M,Dd)).astype(np.float32)
    m.mlp.get_output(dev(ep),dev(ed),1,R)
    cg,dg=rng.normal(size=(M,3)).astype(np.float32),rng.normal(size=M).astype(np.float32)
    m.mlp.reset_gradients(1); m.mlp.get_gradient(dev(cg),dev(dg),1)
    rd,rr,acts=orc.mlp_forward(ocfg,params,ep,ed,prec='f64')
    d_rd,d_rr=orc.output_activations_grad(ocfg,rd,rr,dg,cg,prec='f64')
    g64=orc.mlp_backward(ocfg,params,ep,ed,acts,d_rd,d_rr,prec='f64')
    got=m.get_gradients(); sizes=m.GetLayerSizes(); off=0; errs=[]
    for i,n in enumerate(sizes):
        errs.append(rel_err(got[off:off+n],g64[off:off+n])); off+=n
    print(precision,'M',M,'total',f'{rel_err(got,g64):.2e}',' '.join(f'{e:.1e}' for e in errs))
NET=dict(n_samples=64)
for prec in ('fp32_tc','bf16','fp32'):
    for R in (3,4,8,16,64):
        run(prec,R,NET)
