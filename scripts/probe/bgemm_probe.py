import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import rel_l2
from ocrs_models_b200 import _lib
from ocrs_models_b200._lib import call, ptr
lib = _lib.lib(); st = _lib.stream_ptr(torch.device("cuda:0"))
for (N, cin, cout, HW) in [(4, 32, 64, 4096), (4, 32, 64, 1024), (1, 32, 64, 4096), (4, 32, 128, 4096), (4, 64, 64, 4096)]:
    g = torch.Generator().manual_seed(N * cin + HW)
    x = torch.randn(N, cin, HW, generator=g); dy = torch.randn(N, cout, HW, generator=g)
    xd, dyd = x.cuda(), dy.cuda()
    part = torch.empty((N, cout, cin), device="cuda")
    call("ocrs_gemm_tc_batched", ptr(dyd), HW, 1, N * cout, cout, ptr(xd), HW, 1, N * cin, cin, ptr(part), cin, cout * cin, cout, cin, HW, N, None, st)
    ref = torch.einsum("nop,nip->noi", dy.double(), x.double())
    print((N, cin, cout, HW), "total", rel_l2(part.double().sum(0), ref.sum(0)), "per item", [round(rel_l2(part[n], ref[n]), 8) for n in range(N)])
