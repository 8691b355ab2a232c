"""Torch-tensor front ends of the CUDA entry points (device pointers + current stream only;
all arithmetic happens in libasd_b200.so)."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from ._lib import AsdError, check, lib

NUM_FEATURES = 6
FEATURE_NAMES = ("lse", "p_max", "margin", "entropy", "logprob_draft_token", "logprob_resampled_token")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class RejectionSampler:
    """Owns the zero-initialised workspace of ``asd_reject_sample`` for fixed (B, k)."""

    def __init__(self, B: int, k: int, device="cuda"):
        self.B, self.k = B, k
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        nbytes = lib().asd_reject_sample_workspace_bytes(B, k)
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self.accept_mask = torch.empty(B, k, dtype=torch.uint8, device=self.device)
        self.accepted_len = torch.empty(B, dtype=torch.int32, device=self.device)
        self.out_tokens = torch.empty(B, k + 1, dtype=torch.int32, device=self.device)
        self.out_logprobs = torch.empty(B, k + 1, dtype=torch.float32, device=self.device)
        self.features = torch.empty(B, k + 1, NUM_FEATURES, dtype=torch.float32, device=self.device)

    def __call__(self, target_logits: torch.Tensor, draft_logits: Optional[torch.Tensor],
                 draft_tokens: torch.Tensor, u_accept: torch.Tensor, u_resid: torch.Tensor, temperature: float):
        B, k = self.B, self.k
        if not target_logits.is_cuda:
            raise AsdError("asd_reject_sample needs CUDA tensors (no CPU fallback)")
        V = target_logits.shape[-1]
        assert target_logits.dtype == torch.float32 and target_logits.is_contiguous()
        assert target_logits.numel() == B * (k + 1) * V
        if draft_logits is not None:
            assert draft_logits.dtype == torch.float32 and draft_logits.is_contiguous()
            assert draft_logits.numel() == B * k * V
        assert draft_tokens.dtype == torch.int32 and draft_tokens.numel() == B * k
        assert u_accept.dtype == torch.float64 and u_accept.numel() == B * k
        assert u_resid.dtype == torch.float64 and u_resid.numel() == B
        with torch.cuda.device(self.device):
            rc = lib().asd_reject_sample(
                target_logits.data_ptr(), 0 if draft_logits is None else draft_logits.data_ptr(),
                draft_tokens.data_ptr(), u_accept.data_ptr(), u_resid.data_ptr(), B, k, V, float(temperature),
                self.accept_mask.data_ptr(), self.accepted_len.data_ptr(), self.out_tokens.data_ptr(),
                self.out_logprobs.data_ptr(), self.features.data_ptr(), self.workspace.data_ptr(), _stream())
        check(rc, "asd_reject_sample")
        return dict(accept_mask=self.accept_mask, accepted_len=self.accepted_len, out_tokens=self.out_tokens,
                    out_logprobs=self.out_logprobs, features=self.features)


def reject_sample_host(target_logits, draft_logits, draft_tokens, u_accept, u_resid, temperature):
    """numpy (HOST) buffers in and out through ``asd_reject_sample_host``: the call a non-torch
    integrator would make; copies happen inside the library."""
    import numpy as np
    tl = np.ascontiguousarray(target_logits, dtype=np.float32)
    B, k1, V = tl.shape
    k = k1 - 1
    dl = None if draft_logits is None else np.ascontiguousarray(draft_logits, dtype=np.float32)
    dt = np.ascontiguousarray(draft_tokens, dtype=np.int32).reshape(B, k)
    ua = np.ascontiguousarray(u_accept, dtype=np.float64).reshape(B, k)
    ur = np.ascontiguousarray(u_resid, dtype=np.float64).reshape(B)
    out = dict(accept_mask=np.zeros((B, k), np.uint8), accepted_len=np.zeros(B, np.int32),
               out_tokens=np.zeros((B, k + 1), np.int32), out_logprobs=np.zeros((B, k + 1), np.float32),
               features=np.zeros((B, k + 1, NUM_FEATURES), np.float32))
    rc = lib().asd_reject_sample_host(tl.ctypes.data, None if dl is None else dl.ctypes.data, dt.ctypes.data,
                                      ua.ctypes.data, ur.ctypes.data, B, k, V, float(temperature),
                                      out["accept_mask"].ctypes.data, out["accepted_len"].ctypes.data,
                                      out["out_tokens"].ctypes.data, out["out_logprobs"].ctypes.data,
                                      out["features"].ctypes.data)
    check(rc, "asd_reject_sample_host")
    return out


def linear_bf16(x: torch.Tensor, w: torch.Tensor, out_mode: int = 0, ksplit: int = 0, stages: int = 0):
    """Y = X @ W^T through ``asd_linear_bf16`` (tcgen05/TMA).  out_mode 0 -> fp32 [M, N] (the K-split
    slices are summed here, in order, as the fused consumer kernels do); 1 -> bf16 [M, N];
    2 -> SwiGLU over gate|up-interleaved rows, bf16 [M, N/2]; 3 -> fp32 [M, N] reduced inside the cluster."""
    assert x.is_cuda and w.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    x, w = x.contiguous(), w.contiguous()
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K
    used = ctypes.c_int(0)
    if out_mode == 3:
        out = torch.empty(M, N, dtype=torch.float32, device=x.device)
    elif out_mode == 0:
        ks, st, tt = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        check(lib().asd_linear_plan(M, N, K, 0, ctypes.byref(ks), ctypes.byref(st), ctypes.byref(tt)), "plan")
        nsl = ksplit if ksplit > 0 else ks.value
        out = torch.empty(nsl, M, N, dtype=torch.float32, device=x.device)
    elif out_mode == 1:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=x.device)
    else:
        out = torch.empty(M, N // 2, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib().asd_linear_bf16(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, out_mode, ksplit, stages,
                                   ctypes.byref(used), _stream())
    check(rc, "asd_linear_bf16")
    if out_mode == 0:
        assert used.value == out.shape[0]
        acc = out[0].clone()
        for s in range(1, out.shape[0]):
            acc += out[s]
        return acc
    return out


def linear_bf16_tc(x: torch.Tensor, w: torch.Tensor, out_mode: int = 0, ksplit: int = 0, stages: int = 0):
    """Y = X @ W^T on the tensor-bound CTA-pair kernel (``asd_linear_bf16_tc``).  out_mode 0 -> fp32 [M, N] (the
    K-split slices are summed here in order, as the glue kernels do); 2 -> SwiGLU bf16 [M, N/2]."""
    assert x.is_cuda and w.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    x, w = x.contiguous(), w.contiguous()
    M, K = x.shape
    N = w.shape[0]
    used = ctypes.c_int(0)
    if out_mode == 2:
        out = torch.empty(M, N // 2, dtype=torch.bfloat16, device=x.device)
    else:
        out = torch.empty(8, M, N, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib().asd_linear_bf16_tc(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, out_mode, ksplit, stages,
                                      ctypes.byref(used), _stream())
    check(rc, "asd_linear_bf16_tc")
    if out_mode == 2:
        return out
    acc = out[0].clone()
    for s in range(1, used.value):
        acc += out[s]
    return acc


def interleave_gate_up(gate_w: torch.Tensor, up_w: torch.Tensor) -> torch.Tensor:
    """[ff, K] gate and up weights -> [2*ceil(ff/64)*64, K] with 64 gate rows then 64 up rows per
    128-row tile (zero rows pad ff to a multiple of 64): the layout asd_linear_bf16 mode 2 expects."""
    ff, K = gate_w.shape
    ffp = (ff + 63) // 64 * 64
    g = torch.zeros(ffp, K, dtype=gate_w.dtype, device=gate_w.device)
    u = torch.zeros(ffp, K, dtype=up_w.dtype, device=up_w.device)
    g[:ff], u[:ff] = gate_w, up_w
    return torch.stack([g.view(ffp // 64, 64, K), u.view(ffp // 64, 64, K)], dim=1).reshape(2 * ffp, K).contiguous()


def cascade_decide(features: torch.Tensor, n_tokens: torch.Tensor, scalars, predictor, prev_p, costs, stage_idx: int,
                   lam: float, prefix_mode: bool = False, risk_adjustment: bool = False, n_obs: float = 100.0,
                   alpha: float = 1.0, beta: float = 1.0):
    """Scorer -> stop decision on the device through ``asd_cascade_decide``.

    features fp32 CUDA [n, T, 6] (the sampler's per-token rows, never copied to the host), n_tokens int32 CUDA [n];
    scalars [n, 3] (prompt words / 2048, output words / 512, stage / 4); predictor: a ``QualityPredictor`` (its MLP
    weights are uploaded once per device and cached on the module); prev_p [n, stage_idx] acceptance probabilities of
    the earlier stages; costs [L].  Returns numpy (prob float64 [n], stop bool [n], k_star int32 [n]) - the only
    device-to-host traffic of the decision."""
    import numpy as np
    if not features.is_cuda:
        raise AsdError("asd_cascade_decide needs CUDA tensors (no CPU fallback)")
    dev = features.device
    n, T, F = features.shape
    assert F == NUM_FEATURES and features.dtype == torch.float32 and features.is_contiguous()
    assert n_tokens.dtype == torch.int32 and n_tokens.numel() == n and n_tokens.device == dev
    L = len(costs)
    cache = getattr(predictor, "_asd_device_weights", None)
    if cache is None:
        cache = predictor._asd_device_weights = {}
    key = (str(dev), predictor.mlp[0].weight._version, predictor.mlp[3].weight._version)
    if key not in cache:
        cache.clear()
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        cache[key] = (f32(predictor.mlp[0].weight), f32(predictor.mlp[0].bias), f32(predictor.mlp[3].weight).view(-1),
                      f32(predictor.mlp[3].bias))
    w1, b1, w2, b2 = cache[key]
    pp = np.ones((n, L), np.float64)
    if stage_idx > 0:
        pp[:, :stage_idx] = np.asarray(prev_p, dtype=np.float64).reshape(n, stage_idx)
    host = torch.from_numpy(np.concatenate([np.asarray(scalars, np.float64).reshape(n * 3), pp.reshape(-1),
                                            np.asarray(costs, np.float64)]))
    d = host.to(dev, non_blocking=True)
    sc, pv, C = d[:n * 3], d[n * 3:n * 3 + n * L], d[n * 3 + n * L:]
    prob = torch.empty(n, dtype=torch.float64, device=dev)
    out = torch.empty(2, n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib().asd_cascade_decide(features.data_ptr(), n_tokens.data_ptr(), n, T, sc.data_ptr(), w1.data_ptr(),
                                      b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), w1.shape[1], pv.data_ptr(), C.data_ptr(),
                                      L, int(stage_idx), int(bool(prefix_mode)), float(lam), int(bool(risk_adjustment)),
                                      float(n_obs), float(alpha), float(beta), prob.data_ptr(), out[0].data_ptr(),
                                      out[1].data_ptr(), _stream())
    check(rc, "asd_cascade_decide")
    return prob.cpu().numpy(), out[0].cpu().numpy().astype(bool), out[1].cpu().numpy()
