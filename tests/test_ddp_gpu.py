"""Multi-GPU data parallelism on hardware: with the real kernels and NCCL, the gradient in the
flat bucket after the all-reduce equals the mean of the per-shard fp64-oracle gradients, and a 2-rank FusedAdam step
leaves both replicas with identical weights. Skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


def _shard(kind, rank):
    g = torch.Generator().manual_seed(1234 + rank)  # bench.make_batch's per-rank seed
    if kind == "rec":
        return {"image": torch.rand(3, 1, 64, 128, generator=g) - 0.5, "targets": torch.randint(1, 97, (3, 16), generator=g, dtype=torch.int32),
                "input_lengths": torch.full((3,), 32, dtype=torch.int64), "target_lengths": torch.tensor([16, 9, 2])}
    return {"image": torch.rand(2, 1, 96, 64, generator=g) - 0.5, "mask": (torch.rand(2, 1, 96, 64, generator=g) < 0.1).float()}


def _worker(rank, world, port, kind, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    torch.distributed.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from ocrs_models_b200 import CTCLoss, DetectionModel, RecognitionModel, balanced_cross_entropy_loss
    from ocrs_models_b200.alphabet import DEFAULT_ALPHABET
    from ocrs_models_b200.optim import FusedAdam
    from oracle import functional as O

    torch.manual_seed(1234)
    model = RecognitionModel(DEFAULT_ALPHABET) if kind == "rec" else DetectionModel()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(dev).train()
    opt = FusedAdam(model, lr=1e-3, max_grad_norm=4.0 if kind == "rec" else None, world_size=world)
    b = _shard(kind, rank)
    opt.zero_grad()
    if kind == "rec":
        loss = CTCLoss()(model(b["image"].to(dev)), b["targets"].to(dev), b["input_lengths"], b["target_lengths"])
    else:
        loss = balanced_cross_entropy_loss(model(b["image"].to(dev)), b["mask"].to(dev))
    loss.backward()
    avg = {k: g.cpu().double() for (k, _), (_, g) in zip(model.named_parameters(), opt.averaged_gradients())}
    # mean of the per-shard oracle gradients (every rank recomputes both shards on the CPU: small shapes)
    ref = None
    for r in range(world):
        g = O.train_step_grads(kind, sd, _shard(kind, r), torch.float64)[2]
        ref = g if ref is None else {k: ref[k] + g[k] for k in g}
    ref = {k: v / world for k, v in ref.items()}
    gn = torch.sqrt(sum((v ** 2).sum() for v in ref.values()))
    err = float(torch.sqrt(sum(((avg[k] - ref[k]) ** 2).sum() for k in ref)) / gn)
    opt.step()
    torch.cuda.synchronize()
    flat = opt.flat_p.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    torch.distributed.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered[1:])
    out.put((rank, err, bool(same)))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("kind", ["rec", "det"])
def test_allreduced_gradient_is_the_mean_of_shard_oracle_gradients(kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200 + (0 if kind == "rec" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    print(f"{kind}: averaged-gradient rel-L2 vs mean of shard oracles: {[r[1] for r in res]}")
    for rank, err, same in res:
        assert err < 1e-3, (rank, err)
        assert same, "replicas diverged after the step"
