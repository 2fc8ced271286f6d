import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("clip", [None, 4.0, 0.05])
def test_fused_adam_matches_torch(clip):
    from ocrs_models_b200.optim import FusedAdam

    torch.manual_seed(0)
    m1 = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)).cuda()
    m2 = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5)).cuda()
    m2.load_state_dict(m1.state_dict())
    ref = torch.optim.Adam(m1.parameters(), lr=1e-3)
    ours = FusedAdam(m2, lr=1e-3, max_grad_norm=clip)
    for it in range(4):
        x = torch.randn(8, 37, device="cuda")
        for m, opt in ((m1, ref), (m2, ours)):
            opt.zero_grad()
            (m(x) ** 2).sum().backward()
        n_ref = torch.nn.utils.clip_grad_norm_(m1.parameters(), clip) if clip else None
        ref.step()
        n = ours.step()
        if clip:
            assert abs(n.item() - n_ref.item()) < 1e-4 * n_ref.item()
        for a, b in zip(m1.parameters(), m2.parameters()):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=1e-7)
