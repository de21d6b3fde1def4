cd $GRAFT_REPO_ROOT
python - <<'PY'
import torch
from seald_nerf_b200 import _lib
from seald_nerf_b200._lib import ptr
dev=torch.device('cuda:0')
out=torch.zeros(148*2,dtype=torch.int64,device=dev)
for n_cols, modes in ((128,(0,1)), (64,(0,1,2,3)), (32,(0,))):
  for mode in modes:
    for issuers in (1,2,4):
        iters=512
        _lib.call("seald_umma_probe", mode, iters, n_cols, issuers, ptr(out), _lib.stream())
        torch.cuda.synchronize()
        o=out.view(148,2).float()
        tot=iters*issuers
        print("N", n_cols, "mode",mode,"issuers",issuers,"cyc per mma (all issuers) %.1f   per-issuer issue cyc/mma %.1f"%(o[:,1].mean()/tot, o[:,0].mean()/iters), flush=True)
PY
bash scripts/gpu/t.sh
