"""ctypes binding of the C-ABI library (include/ocrs_b200.h).

The product path has no CPU or PyTorch fallback: if ``libocrs_b200.so`` is missing, or a call
returns non-zero, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libocrs_b200.so")

P, I, F, D, L = c_void_p, c_int, c_float, c_double, c_longlong

# name -> argtypes (all return int unless listed in _RESTYPE)
SIGNATURES: dict[str, list] = {
    "ocrs_version": [],
    "ocrs_device_arch": [],
    "ocrs_launch_count": [],
    "ocrs_ctc_alpha_row": [I],
    "ocrs_ctc_fwd": [P, P, I, P, P, I, I, I, I, I, I, I, P, P, P, P],
    "ocrs_ctc_bwd": [P, P, I, P, P, I, I, I, I, I, I, I, P, P, P, P, P],
    # detection forward (csrc/det_fwd.cu)
    "ocrs_det_dwpw_partial_rows": [I, I, I],
    "ocrs_det_dwpw_fwd": [P, L, I, I, I, I, P, P, P, P, P, I, P, L, P, P],
    "ocrs_det_dw3x3_fwd": [P, L, I, I, I, I, P, P, P, P, P, P],
    "ocrs_det_activate": [P, L, I, I, L, P, P, P, P, P],
    "ocrs_det_convt_col2im": [P, I, I, I, I, P, P, L, I, I, P],
    "ocrs_bn_finalize": [P, I, I, D, P, P, P, P, F, F, I, I, P, P, P, P, P, P, P],
    "ocrs_det_pool2_fwd": [P, L, I, I, I, I, P, P, P, P, L, P],
    "ocrs_det_convt_fwd": [P, L, I, I, I, I, P, P, P, P, P, I, P, L, I, I, P],
    "ocrs_det_outconv_fwd": [P, L, I, I, I, I, P, P, P, P, P, P, P],
    # detection backward (csrc/det_bwd.cu)
    "ocrs_finalize_partials": [P, I, I, P, P],
    "ocrs_det_outconv_bwd_rows": [I, I, I],
    "ocrs_det_outconv_bwd": [P, P, P, L, I, I, I, I, P, P, P, P, P, L, P, P],
    "ocrs_reduce_rows": [I, L],
    "ocrs_bnrelu_bwd_reduce": [P, L, P, L, I, I, L, P, P, P, P, P, P, P],
    "ocrs_bn_bwd_finalize": [P, I, I, D, P, P, P, P, P, P, P, P, I, P],
    "ocrs_det_dy": [P, L, P, L, I, I, L, P, P, P, P, P, P, P, P],
    "ocrs_det_convt_im2col": [P, L, I, I, I, I, I, I, P, P],
    "ocrs_det_pwT_bwd": [P, L, P, L, I, I, L, P, P, P, P, P, P, P, I, P, L, P],
    "ocrs_det_pw_wgrad_workers": [I, I, I],
    "ocrs_det_pw_wgrad": [P, L, P, L, I, I, I, I, P, P, P, P, P, P, P, L, I, P, P, P, P, P, P],
    "ocrs_det_dw_bwd_rows": [I, I, I],
    "ocrs_det_dw_bwd": [P, L, P, L, I, I, I, I, P, P, P, P, P, L, I, P, P],
    "ocrs_det_pool2_bwd": [P, L, I, I, I, I, P, P, P, P, L, P, L, P],
    "ocrs_det_pool2_bwd_bn_rows": [I, I, I],
    "ocrs_det_pool2_bwd_bn": [P, L, I, I, I, I, P, P, P, P, L, P, L, P, P, P, P],
    "ocrs_det_outconv8_bwd_blocks": [],
    "ocrs_det_outconv8_bwd_bn": [P, P, P, L, I, L, P, P, P, P, P, L, P, P, P, P, P],
    "ocrs_det_convt_bwd_data": [P, L, I, I, I, I, P, I, I, I, P, L, P],
    "ocrs_det_convt_wgrad_workers": [I, I, I],
    "ocrs_det_convt_wgrad": [P, L, I, I, I, I, P, P, P, P, L, I, I, I, P, P],
    "ocrs_plane_sum": [P, L, I, I, L, P, P],
    # TMA-pipelined DepthwiseConv kernels (csrc/det_tma.cu)
    "ocrs_det_tma_supported": [P, L, P, L, I, I],
    "ocrs_det_sep_channels_ok": [I, I],
    "ocrs_det_sep_fwd_rows": [I, I, I, I],
    "ocrs_det_sep_fwd": [P, L, I, I, I, I, P, P, P, P, P, I, P, L, P, P, P],
    "ocrs_det_pw_wgrad_saved_workers": [I, L, I, I],
    "ocrs_det_pw_wgrad_saved": [P, L, P, L, I, I, L, P, P, P, P, P, P, P, I, P, P, P, L, P],
    "ocrs_det_pw_wgrad_saved32_workers": [I, L, I],
    "ocrs_det_pw_wgrad_saved32": [P, L, P, L, I, L, P, P, P, P, P, P, P, I, P, P, P, L, P],
    "ocrs_det_sep_dw_bwd_rows": [I, I, I, I],
    "ocrs_det_sep_dw_bwd": [P, L, P, L, I, I, I, I, P, P, P, P, P, L, I, P, P, P, P, P],
    "ocrs_det_sep_pw_wgrad_workers": [I, I, I, I, I],
    "ocrs_det_sep_pw_wgrad": [P, L, P, L, I, I, I, I, P, P, P, P, P, P, P, L, I, P, P, P, P, P, P],
    "ocrs_det_convt_wgrad_staged_ok": [P, L, I, I, P, L, I, I],
    "ocrs_det_convt_wgrad_staged_workers": [I, I, I, I, I],
    "ocrs_det_convt_wgrad_staged": [P, L, I, I, I, I, P, P, P, P, L, I, I, I, P, P],
    # GEMM / im2col (csrc/gemm.cu)
    "ocrs_gemm_stat_rows": [I],
    "ocrs_gemm": [P, L, I, P, L, I, P, L, I, I, I, P, I, I, P, I, P],
    "ocrs_gemm_splits": [I, I],
    # tcgen05 GEMM (csrc/gemm_tc.cu)
    "ocrs_gemm_tc_set_fast": [I],
    "ocrs_gemm_tc_supported": [P, L, P, L],
    "ocrs_gemm_tc": [P, L, I, P, L, I, P, L, I, I, I, P, I, I, P, I, P],
    "ocrs_gemm_tc_splits": [I, I],
    "ocrs_gemm_tc_batched": [P, L, I, I, I, P, L, I, I, I, P, L, L, I, I, I, I, P, P],
    "ocrs_gemm_tc_batched_splits": [I, I],
    "ocrs_gemm_tc_batched_splitk": [P, L, I, I, I, P, L, I, I, I, P, L, L, I, I, I, I, I, P],
    "ocrs_gemm_tc_batched_stat_rows": [I, I],
    "ocrs_gemm_tc_presplit": [P, L, I, P, P, L, I, P, L, I, I, I, P, I, I, P, I, P],
    "ocrs_conv3x3_tc_presplit": [P, I, I, I, I, P, P, I, P, L, P, I, P, P],
    "ocrs_split_tf32": [P, P, P, L, P],
    "ocrs_conv3x3_tc": [P, I, I, I, I, P, I, P, L, P, I, P, P],
    "ocrs_conv3x3_wgrad_tc": [P, P, I, I, I, I, I, P, I, P],
    "ocrs_im2col_nhwc": [P, I, I, I, I, I, I, I, I, I, I, P, P],
    "ocrs_colsum_rows": [I],
    "ocrs_colsum": [P, L, I, I, P, P],
    # recognition (csrc/rec.cu)
    "ocrs_rec_conv0_fwd": [P, I, I, I, P, P, P, P, P],
    "ocrs_rec_conv0_bwd_blocks": [],
    "ocrs_rec_conv0_bwd": [P, I, I, I, P, P, P, P, P, P],
    "ocrs_rec_bn_act_pool_fwd": [P, I, I, I, I, I, I, I, I, P, P, P, L, L, L, P],
    "ocrs_rec_pool_bwd_blocks": [],
    "ocrs_rec_bn_act_pool_bwd_reduce": [P, I, I, I, I, I, I, I, I, P, P, P, P, P, L, L, L, P, P],
    "ocrs_rec_bn_act_pool_bwd_apply": [P, I, I, I, I, I, I, I, I, P, P, P, P, P, P, L, L, L, P, P],
    "ocrs_relu_bwd": [P, P, L, P],
    "ocrs_gru_layer_fwd_persist": [P, P, P, P, P, P, P, P, I, I, P],
    "ocrs_gru_layer_bwd_persist": [P, P, P, P, P, P, P, P, P, I, I, P],
    "ocrs_log_softmax_fwd": [P, P, I, I, P],
    "ocrs_log_softmax_bwd": [P, P, P, I, I, I, P],
    "ocrs_transpose": [P, P, I, I, P],
    # recognition accuracy bookkeeping (csrc/metrics.cu)
    "ocrs_ctc_greedy_cer_max_targets": [],
    "ocrs_ctc_greedy_cer": [P, I, I, I, P, P, L, I, I, P, P, P, P, P],
    # input pipeline (csrc/data.cu)
    "ocrs_collate_lines": [P, I, P, P, I, I, I, P, P],
    # eval-path post-processing (csrc/postprocess.cu)
    "ocrs_cc_label": [P, F, I, I, I, P, P, P, P],
    "ocrs_cc_boundary": [P, I, I, I, P, I, P, P],
    # multi-tensor step glue (csrc/glue.cu)
    "ocrs_grad_deliver_max": [],
    "ocrs_weight_prep_max": [],
    "ocrs_grad_deliver": [P, P, P, P, P, P, P, P, P, I, P],
    "ocrs_weight_prep": [P, P, P, P, P, P, P, P, P, P, P, I, P],
    # optimiser glue (csrc/optim.cu)
    "ocrs_optim_blocks": [],
    "ocrs_grad_norm": [P, L, F, P, P, P],
    "ocrs_adam_step": [P, P, P, P, L, F, F, F, F, I, F, F, P, P],
    "ocrs_adam_step_dev": [P, P, P, P, L, F, F, F, F, P, F, F, P, P],
    # balanced BCE (csrc/det_loss.cu)
    "ocrs_bce_state_words": [],
    "ocrs_bce_blocks": [],
    "ocrs_balanced_bce_fwd": [P, P, L, P, P, P, P, P],
    "ocrs_balanced_bce_bwd": [P, P, P, L, P, P, P, P],
}
_RESTYPE = {"ocrs_last_error": c_char_p, "ocrs_launch_count": c_longlong}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libocrs_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: run `make -C {CSRC}` (or __graft_entry__.build()). "
                "There is no CPU fallback for this path."
            )
        l = ctypes.CDLL(LIB_PATH)
        l.ocrs_last_error.restype = c_char_p
        l.ocrs_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, c_int)
        _lib = l
    return _lib


# Optional per-entry-point device timing (bench.py's roofline pass): name -> [(start, end, meta)].
PROFILE: dict | None = None


def call(name: str, *args, meta=None) -> None:
    """Invoke an entry point; raise RuntimeError with ocrs_last_error() on failure."""
    l = lib()
    if PROFILE is not None:
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(l, name)(*args)
        e1.record()
        PROFILE.setdefault(name, []).append((e0, e1, meta))
    else:
        rc = getattr(l, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {l.ocrs_last_error().decode()}")


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def check_module_tensors(model, device, what: str) -> None:
    """The kernels take raw pointers: every parameter and buffer must be a contiguous fp32 (int64 for
    num_batches_tracked) CUDA tensor on the input's device. Anything else (model left on the CPU or another GPU,
    .half()/.bfloat16(), channels_last) raises here instead of faulting inside a kernel."""
    import torch

    for kind, items in (("parameter", model.named_parameters()), ("buffer", model.named_buffers())):
        for name, t in items:
            if t.device != device:
                raise RuntimeError(f"{what}: {kind} {name} is on {t.device} but the input is on {device} "
                                   "(there is no CPU path: move the model with .to(device))")
            if t.is_floating_point() and t.dtype != torch.float32:
                raise RuntimeError(f"{what}: {kind} {name} has dtype {t.dtype}; the kernels compute in float32")
            if not t.is_contiguous():
                raise RuntimeError(f"{what}: {kind} {name} is not contiguous (memory_format / strided views are not supported)")


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
