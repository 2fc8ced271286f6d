"""Device-side greedy CTC decode + CER (csrc/metrics.cu) vs the oracle restatement of train_rec.py:29-68: bit exact."""
import pytest
import torch

from oracle import functional as O

pytestmark = pytest.mark.gpu


def _case(T, N, C, S_pad, g, peaky=True):
    lp = torch.randn(T, N, C, generator=g)
    if peaky:  # long runs of repeated labels and blanks, like a trained CTC model
        runs = torch.randint(0, C, (T // 3 + 1, N), generator=g).repeat_interleave(3, dim=0)[:T]
        runs[torch.rand(T, N, generator=g) < 0.4] = 0
        lp.scatter_(2, runs.unsqueeze(-1), 10.0)
    lp = torch.log_softmax(lp, 2)
    tg = torch.randint(1, C, (N, S_pad), generator=g, dtype=torch.int32)
    tl = torch.randint(0, S_pad + 1, (N,), generator=g)
    for n in range(N):
        tg[n, tl[n]:] = 0
    if N > 1 and S_pad > 2:
        tg[1, 0] = 0  # a blank inside the row is skipped by decode_text too
    pl = torch.randint(0, T + 1, (N,), generator=g)
    pl[0] = T
    return lp, pl, tg, tl


@pytest.mark.parametrize("T,N,C,S_pad", [(201, 64, 97, 64), (257, 5, 97, 128), (12, 4, 6, 5), (65, 33, 97, 1), (40, 3, 11, 256), (7, 2, 3, 0)])
def test_greedy_cer_bit_exact(T, N, C, S_pad):
    from ocrs_models_b200 import greedy_decode_cer

    g = torch.Generator().manual_seed(T * 31 + S_pad)
    lp, pl, tg, tl = _case(T, N, C, S_pad, g)
    ref_d, ref_dec = O.greedy_cer(lp, pl, tg)
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    dist, dec, dlen = greedy_decode_cer(lp.cuda(), pl, tg.cuda() if S_pad else tg.reshape(N, 0).cuda(), return_decoded=True, total=total)
    assert dist.cpu().tolist() == ref_d
    assert int(total) == sum(ref_d)
    for n in range(N):
        assert dec[n, : int(dlen[n])].cpu().tolist() == ref_dec[n]


def test_stats_class_matches_the_reference_semantics():
    from ocrs_models_b200 import RecognitionAccuracyStats

    g = torch.Generator().manual_seed(5)
    stats = RecognitionAccuracyStats()
    errs = chars = 0
    for _ in range(3):
        lp, pl, tg, tl = _case(65, 8, 97, 64, g)
        stats.update(tg.cuda(), tl, lp.cuda(), pl)
        d, _ = O.greedy_cer(lp, pl, tg)
        errs += sum(d)
        chars += int(tl.sum())
    assert stats.char_errors == errs and stats.total_chars == chars
    assert stats.stats_dict() == {"char_error_rate": errs / chars}
