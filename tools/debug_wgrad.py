import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contrad_b200 import kernels as K
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
M, N, Kd = 64, 128, 32
torch.manual_seed(0)
dy = K.round_tf32(torch.randn(M, N, device="cuda")); x = K.round_tf32(torch.randn(M, Kd, device="cuda"))
dw = K.gemm_tn_wgrad(dy, x); ref = (dy.double().t() @ x.double()).float()
print("random: dw absmax", float(dw.abs().max()), "ref absmax", float(ref.abs().max()), "nonzero frac", float((dw != 0).float().mean()))
print("dw[:4,:8]\n", dw[:4, :8].cpu(), "\nref[:4,:8]\n", ref[:4, :8].cpu())
r = (dw / ref)
print("ratio median", float(r.median()), "ratio[:2,:6]", r[:2, :6].cpu())
for (m0, n0) in [(0, 0), (1, 0), (0, 1), (8, 0), (9, 33), (63, 127), (5, 31), (5, 32)]:
    dy = torch.zeros(M, N, device="cuda"); dy[m0, n0] = 1.0
    x = (torch.arange(Kd, device="cuda").float()[None, :] + 1 + 100 * torch.arange(M, device="cuda").float()[:, None])
    dw = K.gemm_tn_wgrad(dy, x)
    nz = dw.nonzero()
    print("one-hot dy[%d,%d]: expect row %d = %s.. ; got %d nonzeros; rows %s; first vals %s" % (
        m0, n0, n0, x[m0, :4].tolist(), nz.shape[0], sorted(set(nz[:, 0].tolist()))[:8],
        [(int(a), int(b), float(dw[a, b])) for a, b in nz[:6].tolist()]))
