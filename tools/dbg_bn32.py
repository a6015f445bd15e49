import sys, torch
sys.path.insert(0, '/root/repo')
from tensorlink_b200 import native as nat
torch.manual_seed(0)
for (M,N,K) in [(32,3584,18944),(32,2048,512),(9,4608,3584),(32,64,64),(32,32,64),(128,96,128)]:
    a=(torch.randn(M,K)).bfloat16().cuda(); w=(torch.randn(N,K)*0.05).bfloat16().cuda()
    got=nat.gemm(a,w,flags=nat.EPI_OUT_F32)
    ref=a.float()@w.float().t()
    err=(got-ref).abs()
    rel=float((got-ref).norm()/ref.norm())
    bad=(err>1e-2*ref.abs().max()).nonzero()
    print((M,N,K),'rel',rel,'nbad',len(bad), 'first bad', bad[:5].tolist(), 'bad cols mod 32', sorted(set((bad[:,1]%32).tolist()))[:40] if len(bad) else None, 'bad rows', sorted(set(bad[:,0].tolist()))[:10] if len(bad) else None)
