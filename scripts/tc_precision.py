"""How does the 3xTF32 tcgen05 GEMM error scale with K (accumulation behaviour)?"""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ocrs_models_b200 import rec_engine, _lib
from ocrs_models_b200.rec_engine import gemm
st = _lib.stream_ptr(torch.device("cuda:0"))
g = torch.Generator().manual_seed(0)
for K in (64, 256, 1024, 4096, 16384):
    A = torch.randn(256, K, generator=g); B = torch.randn(128, K, generator=g)
    for name, (a, b) in {"fp32 inputs": (A, B), "tf32-exact inputs": ((A.view(torch.int32) & ~0x1fff).view(torch.float32), (B.view(torch.int32) & ~0x1fff).view(torch.float32)),
                         "positive tf32-exact": ((A.abs().view(torch.int32) & ~0x1fff).view(torch.float32), (B.abs().view(torch.int32) & ~0x1fff).view(torch.float32))}.items():
        ref = a.double() @ b.double().t()
        res = {}
        for be in ("tc", "simt"):
            rec_engine.GEMM_BACKEND = be
            out = gemm(a.cuda(), K, True, b.cuda(), K, True, 256, 128, K, st)
            d = out.cpu().double() - ref
            res[be] = (float(d.norm() / ref.norm()), float(d.mean() / ref.abs().mean()))
        print(f"K={K:6d} {name:22s} tc rel {res['tc'][0]:.2e} bias {res['tc'][1]:+.2e} | simt rel {res['simt'][0]:.2e} bias {res['simt'][1]:+.2e}")
