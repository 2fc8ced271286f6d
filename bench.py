#!/usr/bin/env python
"""Benchmark of the two training hot paths (BASELINE.json): lines/s of the recognition CRNN train
step at 64x800 batch 64 per GPU (headline line) and images/s of the detection train step at
1024x1024 batch 32 per GPU (reported under "det"). One step = fwd + loss + bwd + (clip) + Adam.

    python bench.py --gpus N --steps K --warmup W            # N > 1: launched through torchrun
    python bench.py --impl reference ...                     # the unmodified reference (baseline/_ref) on host cores

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

REC = dict(n=64, h=64, w=800, s_pad=64, s=40, il=200)
DET = dict(n=32, h=1024, w=1024)
REC_FLOP_PER_LINE = 11.17e9  # 714.9 GFLOP / 64 lines (fwd + dgrad + wgrad)
DET_BYTES_PER_IMG_FP32 = 2 * 1.655e9  # SURVEY 8d fused-block model, fp32 storage


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops"], tf_sus=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm)}


def make_batch(kind: str, rank: int, device, n: int | None = None):
    g = torch.Generator().manual_seed(1234 + rank)
    if kind == "rec":
        n = n or REC["n"]
        return {
            "image": torch.rand(n, 1, REC["h"], REC["w"], generator=g) - 0.5,
            "targets": torch.randint(1, 97, (n, REC["s_pad"]), generator=g, dtype=torch.int32),
            "input_lengths": torch.full((n,), REC["il"], dtype=torch.int64),
            "target_lengths": torch.full((n,), REC["s"], dtype=torch.int64),
        }
    n = n or DET["n"]
    return {
        "image": torch.rand(n, 1, DET["h"], DET["w"], generator=g) - 0.5,
        "mask": (torch.rand(n, 1, DET["h"], DET["w"], generator=g) < 0.10).float(),
    }


class Workload:
    """Model + fused optimiser + one synthetic batch (pinned on the host, resident on the device)."""

    def __init__(self, kind, device, rank, world, group=None, n=None):
        from ocrs_models_b200 import CTCLoss, DetectionModel, RecognitionModel, balanced_cross_entropy_loss
        from ocrs_models_b200.optim import FusedAdam
        from ocrs_models_b200.alphabet import DEFAULT_ALPHABET

        self.kind, self.device = kind, device
        torch.manual_seed(1234)  # train_detection.py:337-338
        if kind == "rec":
            self.model = RecognitionModel(DEFAULT_ALPHABET).to(device).train()
            self.opt = FusedAdam(self.model, lr=1e-3, max_grad_norm=4.0, process_group=group, world_size=world)
            self.loss_fn = CTCLoss()
        else:
            self.model = DetectionModel().to(device).train()
            self.opt = FusedAdam(self.model, lr=1e-3, process_group=group, world_size=world)
            self.loss_fn = balanced_cross_entropy_loss
        self.host = {k: v.pin_memory() for k, v in make_batch(kind, rank, device, n).items()}
        self.dev = {k: v.to(device) for k, v in self.host.items()}
        self.units = self.host["image"].shape[0]
        self.h2d_bytes = sum(v.numel() * v.element_size() for k, v in self.host.items() if k in ("image", "mask", "targets"))
        self.graphed = None  # GraphedTrainStep once enable_graph() succeeded

    def enable_graph(self):
        """Capture the whole step into a CUDA graph (ocrs_models_b200.optim.GraphedTrainStep) and replay it from then
        on: the eager step is bound by the host issuing ~280 launches, not by the kernels. Falls back to eager."""
        from ocrs_models_b200.optim import GraphedTrainStep

        def loss_from_batch(model, b):
            if self.kind == "rec":
                return self.loss_fn(model(b["image"]), b["targets"], b["input_lengths"], b["target_lengths"])
            return self.loss_fn(model(b["image"]), b["mask"])

        try:
            self.graphed = GraphedTrainStep(self.model, self.opt, loss_from_batch, self.dev)
        except Exception as e:  # noqa: BLE001 - any capture problem: keep measuring the eager step and say so
            self.graphed = None
            self.graph_error = f"{type(e).__name__}: {e}"[:300]
        return self.graphed is not None

    def step(self, batch):
        self.opt.zero_grad()
        if self.kind == "rec":
            lp = self.model(batch["image"])
            loss = self.loss_fn(lp, batch["targets"], batch["input_lengths"], batch["target_lengths"])
        else:
            loss = self.loss_fn(self.model(batch["image"]), batch["mask"])
        loss.backward()
        self.opt.step()
        return loss

    def step_eager(self):
        return self.step(self.dev)

    def step_resident(self):
        if self.graphed is not None:
            return self.graphed()
        return self.step(self.dev)

    def step_e2e(self):
        """Public-API step from HOST buffers: H2D of the inputs, the step, D2H of the loss."""
        if self.graphed is not None:
            # pipelined like an input loader with prefetch: the H2D copy of the NEXT step's inputs is issued on a copy
            # stream right after this step's replay is launched; every step still consumes one freshly copied host batch
            hb = {k: self.host[k] for k in ("image", "mask", "targets") if k in self.host}
            g = self.graphed
            if not g._has_staged:
                g.prefetch(hb)
            loss = g()
            g.prefetch(hb)
            return float(loss.item())
        b = dict(self.host)
        for k in ("image", "mask", "targets"):
            if k in b:
                b[k] = b[k].to(self.device, non_blocking=True)
        return float(self.step(b).item())


def timed(fn, steps, warmup, device, dist_on):
    for _ in range(warmup):
        fn()
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    wall = (time.perf_counter() - t0) * 1e3
    if dist_on:
        torch.distributed.barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, wall], device=device, dtype=torch.float64)
    if dist_on:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t[0]), float(t[1])


def kernel_profile(wl: Workload, steps=2):
    """Per-entry-point device time of the step, CUDA events on the launching stream."""
    from ocrs_models_b200 import _lib

    _lib.PROFILE = {}
    for _ in range(steps):
        wl.step_eager()
    torch.cuda.synchronize(wl.device)
    prof, _lib.PROFILE = _lib.PROFILE, None
    out = {}
    for name, evs in prof.items():
        ms = sum(a.elapsed_time(b) for a, b, _ in evs)
        meta = sum(m for _, _, m in evs if m is not None)
        out[name] = dict(ms=ms / steps, calls=len(evs) // steps, meta=meta / steps)
    return out


# entry points that launch the same dominant kernel are judged together
KERNEL_OF = {"ocrs_gemm_tc": "gemm_tc_kernel", "ocrs_conv3x3_tc": "gemm_tc_kernel", "ocrs_conv3x3_wgrad_tc": "gemm_tc_kernel",
             "ocrs_gemm_tc_presplit": "gemm_tc_kernel", "ocrs_conv3x3_tc_presplit": "gemm_tc_kernel",
             "ocrs_gemm": "gemm_kernel", "ocrs_det_pw_wgrad": "pw_wgrad_mma_kernel", "ocrs_det_dwpw_fwd": "dwpw_fwd_kernel",
             "ocrs_det_sep_fwd": "sep_fwd_tma_kernel", "ocrs_det_sep_dw_bwd": "sep_dw_bwd_tma_kernel",
             "ocrs_det_pw_wgrad_saved": "pw_wgrad_saved_kernel", "ocrs_det_sep_pw_wgrad": "sep_pw_wgrad_tma_kernel",
             "ocrs_det_dw_bwd": "dw_bwd_kernel", "ocrs_det_pwT_bwd": "pwT_bwd_kernel",
             "ocrs_bnrelu_bwd_reduce": "bnrelu_bwd_reduce_kernel", "ocrs_det_convt_wgrad": "convt_wgrad_mma_kernel",
             "ocrs_gru_layer_fwd_persist": "gru_fwd_persist_kernel", "ocrs_gru_layer_bwd_persist": "gru_bwd_persist_kernel"}


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    rec = json.load(open(path)).get(kernel)
    return rec["dram_bytes_per_launch"] if rec else None


def roofline(kind, prof, pk):
    """Roofline of the dominant kernel of the step: algorithmic FLOPs or bytes per launch (the `meta` each call site
    attaches) / average launch duration from CUDA events on the launching stream."""
    total = sum(v["ms"] for v in prof.values())
    groups: dict = {}
    for name, v in prof.items():
        g = groups.setdefault(KERNEL_OF.get(name, name.replace("ocrs_", "")), dict(ms=0.0, calls=0, meta=0.0))
        g["ms"] += v["ms"]
        g["calls"] += v["calls"]
        g["meta"] += v["meta"]
    name, v = max(groups.items(), key=lambda kv: kv[1]["ms"])
    share = v["ms"] / total if total else 0.0
    common = dict(kernel=name, share_of_step=share, launches_per_step=v["calls"], avg_launch_ms=v["ms"] / max(v["calls"], 1),
                  traffic=measured_traffic(name))
    if name in ("gemm_tc_kernel", "gemm_kernel"):
        ach = v["meta"] / (v["ms"] * 1e-3) / 1e12
        return dict(bound="tensor", achieved=ach, peak=pk["tf_sus"], unit="TFLOP/s", frac=ach / pk["tf_sus"],
                    peak_source=pk["src"] + " bf16 sustained", **common,
                    note="useful fp32-equivalent FLOP/s of the 3xTF32 tcgen05 GEMM (3 TF32 tensor-core products per useful "
                         "product, TF32 at half the bf16 rate: ceiling = peak/6) against the measured bf16 tensor-pipe peak")
    ach = v["meta"] / (v["ms"] * 1e-3) / 1e9 if v["meta"] else None
    return dict(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=(ach / pk["hbm"]) if ach else None,
                peak_source=pk["src"], **common)


def ctc_saturating(device, pk, n=8192, reps=10):
    """The CTC kernels at the HBM-saturating shape of SURVEY 8d (T=201, N=8192, C=97, S=40): algorithmic bytes
    3*T*N*C*4 + 2*N*T*(2S+1)*4 over the CUDA-event time of forward + backward."""
    from ocrs_models_b200 import _lib
    from ocrs_models_b200._lib import call, ptr

    T, C, S = 201, 97, 40
    st = _lib.stream_ptr(device)
    g = torch.Generator(device=device).manual_seed(0)
    lp = torch.log_softmax(torch.randn(T, n, C, device=device, generator=g), 2)
    tg = torch.randint(1, C, (n, 64), device=device, generator=g, dtype=torch.int32)
    il = torch.full((n,), 200, dtype=torch.int32, device=device)
    tl = torch.full((n,), S, dtype=torch.int32, device=device)
    row = _lib.lib().ocrs_ctc_alpha_row(S)
    alpha = torch.empty(n, T, row, device=device)
    nll, loss, go = torch.empty(n, device=device), torch.empty((), device=device), torch.ones((), device=device)
    grad = torch.empty_like(lp)

    def fwd():
        call("ocrs_ctc_fwd", ptr(lp), ptr(tg), 64, ptr(il), ptr(tl), T, n, C, S, 0, 1, 0, ptr(alpha), ptr(nll), ptr(loss), st)

    def bwd():
        call("ocrs_ctc_bwd", ptr(lp), ptr(tg), 64, ptr(il), ptr(tl), T, n, C, S, 0, 1, 0, ptr(alpha), ptr(nll), ptr(go), ptr(grad), st)

    for _ in range(3):
        fwd(), bwd()
    torch.cuda.synchronize(device)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(reps):
        fwd()
    e[1].record()
    for _ in range(reps):
        bwd()
    e[2].record()
    torch.cuda.synchronize(device)
    tf, tb = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
    alg = 3 * T * n * C * 4 + 2 * n * T * (2 * S + 1) * 4
    ach = alg / ((tf + tb) * 1e-3) / 1e9
    tr = [measured_traffic(k) for k in ("ctc_alpha_kernel", "ctc_beta_grad_kernel")]
    return dict(bound="hbm", kernel="ctc_alpha_kernel + ctc_beta_grad_kernel", shape=f"T={T} N={n} C={C} S={S}",
                fwd_ms=tf, bwd_ms=tb, algorithmic_bytes=alg, achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                peak_source=pk["src"], traffic=(sum(tr) if all(tr) else None))


def cpu_reference_step(kind, n, threads, autocast=False):
    """One training step of the reference's CPU path, returned as a closure -> (step, kind_of_arm).

    kind "reference": the UNMODIFIED reference modules staged under baseline/_ref (scripts/stage_reference.py):
    `DetectionModel` + `balanced_cross_entropy_loss` + Adam exactly as the loop body of train_detection.py:87-98, or
    `RecognitionModel` + `torch.nn.CTCLoss` + `clip_grad_norm_(4.0)` + Adam as train_rec.py:110-151 (the host-side
    CER bookkeeping of :123 is not part of the step definition). fp32; `autocast=True` wraps forward+loss
    in CPU autocast(bfloat16) as train_rec.py:118 does. Falls back to the oracle port (kind "port") only if the
    staged reference is absent."""
    torch.set_num_threads(threads)
    batch = make_batch(kind, 0, "cpu", n)
    sys.path.insert(0, ROOT)
    from baseline import ref_loader

    if ref_loader.available():
        ref_loader.load()
        from ocrs_models import models as ref_models

        torch.manual_seed(1234)
        if kind == "rec":
            from ocrs_models.datasets.hiertext import DEFAULT_ALPHABET as ref_alphabet

            model = ref_models.RecognitionModel(alphabet=ref_alphabet).train()
            opt = torch.optim.Adam(model.parameters(), lr=1e-3)
            loss_fn = torch.nn.CTCLoss()

            def step():
                opt.zero_grad()
                with torch.autocast(device_type="cpu", dtype=torch.bfloat16, enabled=autocast):
                    pred = model(batch["image"])
                    loss = loss_fn(pred, batch["targets"], batch["input_lengths"], batch["target_lengths"])
                loss.backward()
                torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=4.0)
                opt.step()
                return float(loss.item())
        else:
            from ocrs_models.train_detection import balanced_cross_entropy_loss as ref_loss

            model = ref_models.DetectionModel().train()
            opt = torch.optim.Adam(model.parameters())

            def step():
                loss = ref_loss(model(batch["image"]), batch["mask"])
                opt.zero_grad()
                loss.backward()
                opt.step()
                return float(loss.item())

        return step, "reference"

    from ocrs_models_b200 import DetectionModel, RecognitionModel
    from ocrs_models_b200.alphabet import DEFAULT_ALPHABET
    from oracle import functional as O

    torch.manual_seed(1234)
    model = RecognitionModel(DEFAULT_ALPHABET) if kind == "rec" else DetectionModel()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    state: dict = {}

    def step():
        _, loss, grads, nb = O.train_step_grads(kind, sd, batch)
        if kind == "rec":
            O.clip_grad_norm(grads, 4.0)
        O.adam_step({k: sd[k] for k in grads}, grads, state)
        sd.update(nb)
        return float(loss)

    return step, "port"


def time_cpu_steps(step, steps, warmup, cap_s):
    """Run `warmup` + `steps` CPU steps, stopping early (never below 1 warm-up + 1 timed) once `cap_s` is spent."""
    t_begin = time.perf_counter()
    w = 0
    while w < max(warmup, 1):
        step()
        w += 1
        if w >= 1 and time.perf_counter() - t_begin > cap_s * 0.3:
            break
    done, t0 = 0, time.perf_counter()
    loss = None
    while done < max(steps, 1):
        loss = step()
        done += 1
        if time.perf_counter() - t_begin > cap_s:
            break
    return (time.perf_counter() - t0) / done, done, w, loss


def cpu_batch(kind):
    """The reference arm runs the same per-GPU batch as ours for recognition (64 lines, 3.6 GB of host RAM); the
    detection batch of 32 at 1024x1024 needs ~65 GB of host RAM for the autograd graph, so the
    CPU arm times whole images at batch 2 and says so."""
    return REC["n"] if kind == "rec" else 2


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    kind = args.workload
    n = cpu_batch(kind)
    step, arm = cpu_reference_step(kind, n, threads)
    dt, steps, warm, loss = time_cpu_steps(step, args.steps, args.warmup, args.cpu_cap)
    value = n / dt
    unit = "lines/s" if kind == "rec" else "images/s"
    shape = "64x800 lines" if kind == "rec" else "1024x1024 images"
    sample = (f"{n} {shape} per step, {steps} timed steps after {warm} warm-up, fp32, "
              + ("unmodified reference modules (baseline/_ref) + stock torch CTCLoss/clip/Adam" if arm == "reference" else "oracle port"))
    line = {
        "impl": "reference", "metric": metric_name(kind), "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(kind), "batch_per_step": n, "host_threads": threads, "last_loss": loss},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": arm, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if kind == "rec" and arm == "reference" and not args.no_secondary:
        # what train_rec.py:118 literally does on a CPU device: forward+loss under autocast(bfloat16)
        step_ac, _ = cpu_reference_step(kind, n, threads, autocast=True)
        dt_ac, k_ac, _, _ = time_cpu_steps(step_ac, 3, 1, 40.0)
        line["cpu_autocast_bf16"] = {"value": n / dt_ac, "unit": unit, "steps": k_ac,
                                     "note": "same step under torch.autocast('cpu', bfloat16) as train_rec.py:118; not the parity numerics"}
    print(json.dumps(line))


def strong_scaling(args, device, rank, world, dist_on, global_batch=512):
    """BASELINE configs[3]: recognition at a FIXED global batch of 512 lines split over the ranks (strong scaling), next to
    the weak-scaling headline (64 lines per GPU). Same step, device-timed, max over ranks."""
    per = global_batch // world
    wl = Workload("rec", device, rank, world, n=per)
    if not args.no_graph:
        wl.enable_graph()
    steps = max(3, min(args.steps, 10))
    ms, _ = timed(wl.step_resident, steps, 3, device, dist_on)
    del wl
    torch.cuda.empty_cache()
    return {"metric": "train lines/sec rec@64x800, global batch 512 (strong scaling)", "global_batch": per * world,
            "per_gpu_batch": per, "value": per * world * steps / (ms * 1e-3), "unit": "lines/s", "ms_per_step": ms / steps,
            "steps": steps, "scaling": "strong"}


def fast_mode(args, device, rank, world, dist_on):
    """The labelled TF32 fast mode of the recognition path (rec_engine.set_precision("tf32")): same workload, ONE plain TF32
    tensor-core product per k-step instead of 3xTF32. Not parity numerics (error measured by tests/test_fast_mode_gpu.py,
    quoted in DESIGN.md); reported beside the headline, never as the headline."""
    from ocrs_models_b200 import rec_engine

    prev = rec_engine.set_precision("tf32")
    try:
        wl = Workload("rec", device, rank, world)
        if not args.no_graph:
            wl.enable_graph()
        steps = max(3, min(args.steps, 10))
        ms, _ = timed(wl.step_resident, steps, 3, device, dist_on)
        units = wl.units * world
        del wl
        torch.cuda.empty_cache()
    finally:
        rec_engine.set_precision(prev)
    return {"metric": "train lines/sec rec@64x800 (fast mode: plain TF32 GEMMs, NOT parity numerics)", "precision_mode": "tf32",
            "value": units * steps / (ms * 1e-3), "unit": "lines/s", "ms_per_step": ms / steps, "steps": steps}


def metric_name(kind):
    return "train lines/sec rec@64x800" if kind == "rec" else "train images/sec det@1024x1024"


def workload_name(kind):
    if kind == "rec":
        return "text-recognition CRNN training, synthetic 64x800 line images, batch 64 per GPU, CTC loss (BASELINE configs[1])"
    return "text-detection segmentation training, synthetic 1024x1024 greyscale, batch 32 per GPU, balanced BCE (BASELINE configs[2])"


def measure(kind, args, device, rank, world, dist_on, pk):
    from ocrs_models_b200 import _lib

    wl = Workload(kind, device, rank, world)
    lib = _lib.lib()
    # The very first step (initial weights, rank-0 batch) is checked against the loss the UNMODIFIED reference computes
    # for the same seeds in fp64 (tests/golden/bench_first_step.json, written by oracle/bench_constants.py).
    first_loss = float(wl.step_eager().item())
    check = None
    if rank == 0:
        cpath = os.path.join(ROOT, "tests", "golden", "bench_first_step.json")
        ref = json.load(open(cpath))[f"{kind}_loss_f64"]
        rel = abs(first_loss - ref) / abs(ref)
        check = {"ours": first_loss, "reference_fp64": ref, "rel_err": rel, "tolerance": 1e-3}
        if not rel < 1e-3:
            raise SystemExit(f"bench.py: first-step {kind} loss {first_loss} differs from the reference's {ref} (rel {rel:.2e})")
    graph_on = (not args.no_graph) and wl.enable_graph()
    for _ in range(max(args.warmup - 1, 0)):
        wl.step_resident()
    torch.cuda.synchronize(device)
    l0 = lib.ocrs_launch_count()
    with ClockSampler(device.index or 0, enabled=not args.no_clocks) as cs:
        ms, _ = timed(wl.step_resident, args.steps, 0, device, dist_on)
    # replayed graph launches are not seen by the library's host-side counter: one replay = launches_per_step kernels
    launches = wl.graphed.launches_per_step * args.steps if graph_on else lib.ocrs_launch_count() - l0
    _, wall = timed(wl.step_e2e, args.steps, 1, device, dist_on)
    units = wl.units * world
    unit = "lines/s" if kind == "rec" else "images/s"
    res = {
        "value": units * args.steps / (ms * 1e-3), "unit": unit, "ms_per_step": ms / args.steps,
        "e2e": {"value": units * args.steps / (wall * 1e-3), "unit": unit, "h2d_bytes_per_step": wl.h2d_bytes,
                "d2h_bytes_per_step": 4,
                # with the graphed step the H2D copy of step i+1's host batch runs on a copy stream during step i (one copy per
                # step, all inside the timed region), like a data loader with prefetch; the loss read-back syncs every step
                "input_prefetch": bool(graph_on)},
        "gpu_launches": int(launches), "clocks": cs.summary(), "first_step_loss": check,
        "cuda_graph": bool(graph_on) if graph_on else {"enabled": False, "why": getattr(wl, "graph_error", "--no-graph")},
    }
    prof = kernel_profile(wl)  # every rank: the step contains the gradient all-reduce
    if rank == 0:
        res["roofline"] = roofline(kind, prof, pk)
        tot = sum(v["ms"] for v in prof.values())
        res["kernel_ms_per_step"] = {k: round(v["ms"], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]}
        res["kernel_ms_total"] = round(tot, 3)
        if kind == "rec":
            res["ctc_roofline"] = None  # filled below, after the workload's memory is released
            ach = REC_FLOP_PER_LINE * wl.units / (ms / args.steps * 1e-3) / 1e12
            res["step_roofline"] = dict(bound="tensor", achieved=ach, peak=pk["tf_sus"], unit="TFLOP/s", frac=ach / pk["tf_sus"])
        else:
            ach = DET_BYTES_PER_IMG_FP32 * wl.units / (ms / args.steps * 1e-3) / 1e9
            res["step_roofline"] = dict(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                                        note="algorithmic bytes of the ideally fused step at fp32 storage (SURVEY 8d)")
    del wl
    torch.cuda.empty_cache()
    if rank == 0 and kind == "rec":
        res["ctc_roofline"] = ctc_saturating(device, pk)
        torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rec", choices=["rec", "det"], help="headline workload of the JSON line")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other workload's sub-object")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-cap", type=float, default=150.0, help="time cap in seconds of the --impl reference run")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of the CUDA-graph replay")
    ap.add_argument("--no-clocks", action="store_true",
                    help="do not start the nvidia-smi clock sampler (for runs under ncu, which follows child processes)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist_on = world > 1
    if dist_on:
        torch.distributed.init_process_group("nccl", device_id=device)
    pk = peaks()
    args.warmup = max(args.warmup, 3)
    main_res = measure(args.workload, args, device, rank, world, dist_on, pk)
    other = "det" if args.workload == "rec" else "rec"
    other_res = None if args.no_secondary else measure(other, args, device, rank, world, dist_on, pk)

    strong = None if (args.no_secondary or args.workload != "rec") else strong_scaling(args, device, rank, world, dist_on)
    fast = None if (args.no_secondary or args.workload != "rec") else fast_mode(args, device, rank, world, dist_on)
    if rank == 0:
        line = {
            "metric": metric_name(args.workload),
            "value": main_res["value"], "unit": main_res["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "parallelism": f"dp{world}", "l2": "inputs+activations > L2 (126 MB)",
                       "precision_mode": "parity: fp32 storage; 3xTF32 tcgen05 GEMMs (4 TMEM accumulators) + fp32 FMA elsewhere"},
            "e2e": main_res["e2e"], "gpu_launches": main_res["gpu_launches"], "clocks": main_res["clocks"],
            "first_step_loss": main_res["first_step_loss"], "cuda_graph": main_res["cuda_graph"],
            "roofline": main_res.get("roofline"), "step_roofline": main_res.get("step_roofline"),
            "ctc_roofline": main_res.get("ctc_roofline"),
            "kernel_ms_per_step": main_res.get("kernel_ms_per_step"),
        }
        if strong is not None:
            line["strong_scaling"] = strong
        if fast is not None:
            line["fast_mode_tf32"] = fast
        if other_res is not None:
            line[other] = {"metric": metric_name(other),
                           "workload": workload_name(other), **other_res}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n = cpu_batch(args.workload)
            step, arm = cpu_reference_step(args.workload, n, threads)
            dt, reps, warm, _ = time_cpu_steps(step, 5, 1, 25.0)
            line["cpu_baseline"] = {"value": n / dt, "unit": main_res["unit"], "cores": threads, "kind": arm,
                                    "sample": f"{'unmodified reference modules (baseline/_ref)' if arm == 'reference' else 'oracle port'}, "
                                              f"batch {n}, {reps} timed steps after {warm} warm-up, fp32"}
        print(json.dumps(line))
    if dist_on:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
