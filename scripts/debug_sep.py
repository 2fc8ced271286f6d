import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ocrs_models_b200.det_engine import View, _Sep
from ocrs_models_b200.models import _separable
from ocrs_models_b200 import _lib
N, cin, cout, H, W = 2, 1, 8, 37, 45
g = torch.Generator().manual_seed(108)
mod = _separable(cin, cout).cuda()
x = torch.randn(N, cin, H, W, generator=g).cuda()
d_a = torch.randn(N, cout, H, W, generator=g).cuda()
st = _lib.stream_ptr(torch.device('cuda:0'))
prev = None
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    recs = {}
    blk = _Sep(mod)
    y = blk.forward(View(x, 0, cin*H*W, cin, H, W), N, True, st, save=recs)
    grads, dx = blk.backward(recs, View(d_a, 0, cout*H*W, cout, H, W), N, st, None)
    torch.cuda.synchronize()
    cur = [t.clone() for t in grads] + [dx.t.clone(), y.t.clone()]
    if prev is not None:
        print(it, [float((a - b).abs().max()) for a, b in zip(cur, prev)])
    prev = cur
