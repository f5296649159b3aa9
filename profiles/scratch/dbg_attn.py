import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from mmgt_b200.kernels import get_engine
dev = torch.device('cuda', 0)
eng = get_engine(dev, torch.bfloat16)
def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(device=dev, dtype=torch.bfloat16)
for (N, Lq, Lk, Lk2, heads, d) in [(4, 64, 64, 64, 8, 40), (4, 64, 64, 0, 8, 40), (2, 128, 128, 0, 8, 40), (2, 128, 256, 0, 8, 40), (2, 256, 512, 0, 8, 64)]:
    C = heads * d
    q = rnd(N, Lq, 3 * C, seed=51)[:, :, :C]
    kv = rnd(N, Lk, 3 * C, seed=53)
    k, v = kv[:, :, C:2 * C], kv[:, :, 2 * C:]
    for rep in range(3):
        if Lk2:
            bank = rnd(2, Lk2, 2 * C, seed=52)
            idx = torch.tensor([(-1 if i % 3 == 0 else i % 2) for i in range(N)], dtype=torch.int32, device=dev)
            out = eng.attention(q, k, v, heads, k2=bank[:, :, :C], v2=bank[:, :, C:], seg2_index=idx)
        else:
            out = eng.attention(q, k, v, heads)
        torch.cuda.synchronize()
        bad = torch.isnan(out.float()).any(-1)   # (N, Lq)
        print((N, Lq, Lk, Lk2, d), 'rep', rep, 'nan rows:', [(int(a), int(b)) for a, b in bad.nonzero()[:12]], 'count', int(bad.sum()))
