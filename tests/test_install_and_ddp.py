"""Host-side logic that needs no GPU: install() rebinding and the DDP gradient bucket (gloo, world 2)."""
import os
import sys
import textwrap
import types

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_install_rebinds_a_reference_shaped_package(tmp_path, monkeypatch):
    pkg = tmp_path / "fake_ocrs_models"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    (pkg / "models.py").write_text("class DetectionModel: pass\nclass RecognitionModel: pass\n")
    (pkg / "train_detection.py").write_text(textwrap.dedent("""
        from .models import DetectionModel
        def balanced_cross_entropy_loss(p, t): return 'reference'
        def main(): return DetectionModel, balanced_cross_entropy_loss
    """))
    (pkg / "train_rec.py").write_text(textwrap.dedent("""
        from torch.nn import CTCLoss
        from .models import RecognitionModel
        def train(): return CTCLoss, RecognitionModel
    """))
    monkeypatch.syspath_prepend(str(tmp_path))
    import ocrs_models_b200 as ours

    done = ours.install("fake_ocrs_models")
    import fake_ocrs_models.train_detection as td
    import fake_ocrs_models.train_rec as tr

    assert td.main() == (ours.DetectionModel, ours.balanced_cross_entropy_loss)
    assert tr.train() == (ours.CTCLoss, ours.RecognitionModel)
    assert "fake_ocrs_models.models.DetectionModel" in done
    with pytest.raises(ImportError):
        ours.install("no_such_package_anywhere")


@pytest.mark.skipif(not os.path.isdir("/root/reference/ocrs_models"), reason="reference checkout not present (GPU box)")
def test_install_on_the_real_reference_and_its_train_loop_reaches_our_module():
    for name, attrs in {"shapely": [], "shapely.geometry": ["MultiLineString", "JOIN_STYLE", "Polygon"],
                        "shapely.geometry.polygon": ["LinearRing", "Polygon"], "pylev": ["levenshtein"]}.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, object)
            sys.modules[name] = m
    sys.path.insert(0, "/root/reference")
    try:
        import ocrs_models_b200 as ours

        done = ours.install("ocrs_models")
        assert {"ocrs_models.train_detection.balanced_cross_entropy_loss", "ocrs_models.train_rec.CTCLoss"} <= set(done)
        from ocrs_models import train_detection

        model = train_detection.DetectionModel()
        assert isinstance(model, ours.DetectionModel)
        opt = torch.optim.Adam(model.parameters())
        batch = [{"path": ["a"], "image": torch.zeros(1, 1, 64, 64), "text_mask": torch.zeros(1, 1, 64, 64)}]
        # the reference's own train() drives our module; on a CPU-only box it must refuse loudly
        with pytest.raises(RuntimeError, match="no CPU path"):
            train_detection.train(0, torch.device("cpu"), batch, model, train_detection.balanced_cross_entropy_loss, opt)
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "ocrs_models" or k.startswith("ocrs_models.")]:
            del sys.modules[k]


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from ocrs_models_b200.optim import allreduce_sum_, shard_seed

    g = torch.Generator().manual_seed(shard_seed(1234, rank))
    flat = torch.randn(1001, generator=g)
    mine = flat.clone()
    allreduce_sum_(flat)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    torch.distributed.all_gather(gathered, mine)
    ok = torch.allclose(flat / world, torch.stack(gathered).mean(0), atol=1e-6)
    distinct = not torch.equal(gathered[0], gathered[1])
    out.put((rank, bool(ok), bool(distinct)))
    torch.distributed.destroy_process_group()


def test_ddp_flat_bucket_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, True), (1, True, True)]


def test_reference_arm_runs_on_rank0_only():
    import json
    import subprocess

    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
