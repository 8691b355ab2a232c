"""The tensor-bound CTA-pair GEMM (csrc/gemm_tc.cu: tcgen05 cta_group::2, 256 weight rows x <= 256 tokens per tile,
persistent, two TMEM accumulators) against a torch fp32 reference of the same contraction.  bf16 inputs are exact
in fp32, so the only difference is the fp32 summation order: relative tolerance 2e-5 of the row scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w):
    return x.float() @ w.float().T


@pytest.mark.parametrize("M,N,K,ksplit", [
    (576, 1024, 1024, 0),        # 3 token tiles of 192, 4 weight pair-tiles
    (576, 8192, 3696, 0),        # 72B down projection shard at TP = 8 (K not a multiple of 64)
    (300, 1344, 1000, 0),        # ragged everything: 2 token tiles of 160, N not a multiple of 256, K % 64 != 0
    (257, 512, 4096, 3),         # forced 3-way K split: slices summed in order
    (1024, 2560, 8192, 0),       # 4 token tiles of 256: both accumulators full width
    (40, 256, 512, 1),           # smallest legal tile (32 tokens), single unit
])
def test_tc_linear_matches_fp32_reference(M, N, K, ksplit):
    from asd_b200.ops import linear_bf16_tc
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    got = linear_bf16_tc(x, w, 0, ksplit)
    ref = _ref(x, w)
    tol = 2e-5 * ref.abs().max().item() * (K ** 0.5) / 8 + 1e-6
    assert (got - ref).abs().max().item() <= tol, ((got - ref).abs().max().item(), tol)


@pytest.mark.parametrize("M,ff,K", [(576, 3696, 1024), (320, 1000, 768)])
def test_tc_swiglu_matches_reference(M, ff, K):
    from asd_b200.ops import interleave_gate_up, linear_bf16_tc
    g = torch.Generator(device="cuda").manual_seed(ff)
    x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    wg = (torch.randn(ff, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    wu = (torch.randn(ff, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    got = linear_bf16_tc(x, interleave_gate_up(wg, wu), 2)[:, :ff].float()
    ref = torch.nn.functional.silu(_ref(x, wg)) * _ref(x, wu)
    err = (got - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item() + 1e-3, err      # bf16 output rounding


def test_tc_agrees_with_weight_streaming_kernel():
    from asd_b200.ops import linear_bf16, linear_bf16_tc
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randn(576, 2048, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(1536, 2048, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    a, b = linear_bf16_tc(x, w, 0), linear_bf16(x, w, 3)
    assert (a - b).abs().max().item() <= 1e-3 * b.abs().max().item()
