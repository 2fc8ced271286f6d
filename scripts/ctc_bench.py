"""CTC kernel at the HBM-saturating shape of SURVEY 8d (N=8192, T=201, C=97, S=40)."""
import json, os, sys, torch
sys.path.insert(0, '.')
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
from ocrs_models_b200 import _lib
from ocrs_models_b200._lib import call, ptr
dev = torch.device("cuda:0")
st = _lib.stream_ptr(dev)
for N in (64, 1024, 8192):
    T, C, S = 201, 97, 40
    g = torch.Generator(device="cuda").manual_seed(0)
    lp = torch.log_softmax(torch.randn(T, N, C, device=dev, generator=g), 2)
    tg = torch.randint(1, C, (N, 64), device=dev, generator=g, dtype=torch.int32)
    il = torch.full((N,), 200, dtype=torch.int32, device=dev); tl = torch.full((N,), S, dtype=torch.int32, device=dev)
    row = _lib.lib().ocrs_ctc_alpha_row(S)
    alpha = torch.empty(N, T, row, device=dev); nll = torch.empty(N, device=dev); loss = torch.empty((), device=dev)
    grad = torch.empty_like(lp); go = torch.ones((), device=dev)
    def fwd(): call("ocrs_ctc_fwd", ptr(lp), ptr(tg), 64, ptr(il), ptr(tl), T, N, C, S, 0, 1, 0, ptr(alpha), ptr(nll), ptr(loss), st)
    def bwd(): call("ocrs_ctc_bwd", ptr(lp), ptr(tg), 64, ptr(il), ptr(tl), T, N, C, S, 0, 1, 0, ptr(alpha), ptr(nll), ptr(go), ptr(grad), st)
    for _ in range(3): fwd(); bwd()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(10): fwd()
    e[1].record()
    for _ in range(10): bwd()
    e[2].record(); torch.cuda.synchronize()
    tf, tb = e[0].elapsed_time(e[1]) / 10, e[1].elapsed_time(e[2]) / 10
    bytes_alg = 3 * T * N * C * 4 + 2 * N * T * (2 * S + 1) * 4
    print(f"N={N}: fwd {tf:.3f} ms bwd {tb:.3f} ms  algorithmic {bytes_alg/1e9:.3f} GB -> {bytes_alg/1e9/((tf+tb)*1e-3):.0f} GB/s ({bytes_alg/1e9/((tf+tb)*1e-3)/PEAK*100:.1f}% of measured HBM peak)")
