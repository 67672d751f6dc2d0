import sys, torch
sys.path.insert(0, '/root/repo')
from recad_b200 import ops
DEV='cuda:0'
for (M,N,K,bias,relu) in [(1024,512,256,False,False),(1024,1024,512,False,False),(1024,256,128,False,False),(256,512,1024,False,False),(512,1024,1024,False,False),(128,128,32,False,False),(128,128,64,False,False),(128,128,96,False,False),(128,128,128,False,False),(256,128,64,False,False),(128,256,64,False,False),(128,64,64,False,False),(1024,512,1024,True,True),(317,32,64,True,False),(128,64,1000,False,False),(1000,1,64,True,False),(5,200,36,False,True)]:
    g=torch.Generator().manual_seed(M+N+K)
    A,B=torch.randn(M,K,generator=g),torch.randn(N,K,generator=g)
    b=torch.randn(N,generator=g) if bias else None
    C=ops.gemm_tn(A.to(DEV),B.to(DEV),None if b is None else b.to(DEV),relu).cpu().double()
    ref=A.double()@B.double().T+(0 if b is None else b.double())
    if relu: ref=ref.clamp_min(0)
    mag=A.abs().double()@B.abs().double().T+1.0
    err=((C-ref).abs()/mag)
    print((M,N,K,bias,relu), 'max rel err %.3e'%float(err.max()), 'frac bad %.4f'%float((err>2e-6).double().mean()), 'C[0,:3]',C[0,:3].tolist(),'ref',ref[0,:3].tolist())
