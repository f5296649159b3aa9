import sys, torch
sys.path.insert(0, '/root/repo')
from mmgt_b200.kernels import get_engine
dev = torch.device('cuda', 0)
eng = get_engine(dev, torch.bfloat16)
def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(device=dev, dtype=torch.bfloat16)
N, Lq, Lk, heads, d = 2, 256, 512, 8, 64
C = heads * d
q = rnd(N, Lq, 3 * C, seed=51)[:, :, :C]
kv = rnd(N, Lk, 3 * C, seed=53)
out = eng.attention(q, kv[:, :, C:2 * C], kv[:, :, 2 * C:], heads)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
