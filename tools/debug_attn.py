"""dump the first key tile of CTA (0,0,0) of the tensor-core attention kernel (debugging aid)"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import OpRunner, ptr, view
from mug_diffusion_b200.engine import OpList
torch.manual_seed(0)
B, H, D, Lq, Lk = 1, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 32, 48, 48
Cc = H * D
q, k, v = torch.randn(B, Lq, Cc), torch.randn(B, Lk, Cc), torch.randn(B, Lk, Cc)
rel, cg = torch.zeros(129, H), torch.ones(129, H)
R = OpRunner()
qc, kc, vc = q.reshape(-1, Cc).cuda(), k.reshape(-1, Cc).cuda(), v.reshape(-1, Cc).cuda()
out = torch.zeros(B * Lq, Cc).cuda()
dbg = torch.zeros(128 * 40).cuda()
R.lib.mugd_debug_set_attention_dump.argtypes = [C.c_void_p]
R.lib.mugd_debug_set_attention_dump(dbg.data_ptr())
ops = OpList()
relc, cgc = rel.cuda(), cg.cuda()
ops.attention(view(qc), view(kc), view(vc), view(out), ptr(relc), ptr(cgc), B, H, Lq, Lk, 64)
R.run(ops)
torch.cuda.synchronize()
d = dbg.cpu().view(128, 40)
S_ref = (q[0, :, :D] @ k[0, :, :D].t())
print("S raw row0 kernel:", d[0, :8].tolist())
print("S raw row0 ref   :", S_ref[0, :8].tolist())
print("S raw row5 kernel:", d[5, :8].tolist())
print("S raw row5 ref   :", S_ref[5, :8].tolist())
print("K smem rows 0,1 first4:", d[0, 30:34].tolist(), d[1, 30:34].tolist(), " ref:", k[0, 0, :4].tolist(), k[0, 1, :4].tolist())
print("V smem rows 0,1 first4:", d[0, 26:30].tolist(), d[1, 26:30].tolist(), " ref:", v[0, 0, :4].tolist(), v[0, 1, :4].tolist())
P = torch.softmax(S_ref * D ** -0.5, -1)
O_ref = P @ v[0, :, :D]
print("m,l row0:", d[0, 24:26].tolist(), " ref l:", float(torch.exp(S_ref[0] * D ** -0.5 - (S_ref[0] * D ** -0.5).max()).sum()))
print("O tile row0 kernel (unnormalised):", d[0, 16:24].tolist())
print("O row0 ref (normalised)          :", O_ref[0, :8].tolist())
print("out row0:", out[0, :8].tolist())
