"""One training step of a workload between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    ncu --profile-from-start off ... python scripts/one_step.py {rec|det} [batch] [ctc_n]
"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "rec"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
if kind == "ctc":
    # CTC at the HBM-saturating shape of SURVEY 8d
    from ocrs_models_b200 import _lib
    from ocrs_models_b200._lib import call, ptr

    N = n or 8192
    T, C, S = 201, 97, 40
    st = _lib.stream_ptr(dev)
    g = torch.Generator(device="cuda").manual_seed(0)
    lp = torch.log_softmax(torch.randn(T, N, C, device=dev, generator=g), 2)
    tg = torch.randint(1, C, (N, 64), device=dev, generator=g, dtype=torch.int32)
    il = torch.full((N,), 200, dtype=torch.int32, device=dev)
    tl = torch.full((N,), S, dtype=torch.int32, device=dev)
    row = _lib.lib().ocrs_ctc_alpha_row(S)
    alpha = torch.empty(N, T, row, device=dev)
    nll = torch.empty(N, device=dev)
    loss = torch.empty((), device=dev)
    grad = torch.empty_like(lp)
    go = torch.ones((), device=dev)

    def step():
        call("ocrs_ctc_fwd", ptr(lp), ptr(tg), 64, ptr(il), ptr(tl), T, N, C, S, 0, 1, 0, ptr(alpha), ptr(nll), ptr(loss), st)
        call("ocrs_ctc_bwd", ptr(lp), ptr(tg), 64, ptr(il), ptr(tl), T, N, C, S, 0, 1, 0, ptr(alpha), ptr(nll), ptr(go), ptr(grad), st)
else:
    if n:
        (bench.REC if kind == "rec" else bench.DET)["n"] = n
    wl = bench.Workload(kind, dev, 0, 1)
    step = wl.step_resident
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
