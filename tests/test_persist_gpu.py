"""The persistent forward kernel (engine option persist = 1: the whole forward in ONE cooperative launch, stream-K split
of every GEMM over all SMs, phase counters instead of kernel boundaries) against the per-kernel path and the CPU
oracle: same north-star bar (max-abs <= 2e-2), and no device-side wait may have run into its bound."""
import pytest
import torch

from asd_b200.models.qwen2 import Qwen2Config, random_hf_weights, tiny_config
from oracle.model_oracle import qwen2_forward

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,cfg,B,P,q,last", [
    ("tiny-hd64", tiny_config(), 3, 32, 5, False),
    ("tiny-hd64-noprefix", tiny_config(), 2, 0, 7, False),
    ("g5-hd128", Qwen2Config(512, 2, 10, 2, 1024, 4096, head_dim=128, name="g5"), 4, 144, 6, False),
    ("g7-hd128-draft", Qwen2Config(896, 2, 7, 1, 1280, 2048, head_dim=128, name="g7"), 4, 69, 2, True),
])
def test_persistent_forward_matches_oracle_and_per_kernel_path(name, cfg, B, P, q, last):
    from asd_b200.engine import QwenEngine
    w = random_hf_weights(cfg, seed=3, device="cuda", logit_std=0.25)
    ids = torch.randint(0, cfg.vocab_size, (B, P + q), generator=torch.Generator().manual_seed(1))
    idc = ids.cuda().to(torch.int32)
    outs = []
    for persist in (0, 1):
        eng = QwenEngine(cfg, max_seqs=B, max_seq_len=P + q + 16, max_tokens=64).load_hf_weights(w)
        eng.set_option("persist", persist)
        slots = torch.arange(B, dtype=torch.int32, device="cuda")
        if P > 0:
            eng.prefill(idc[:, :P], slots, want_logits=False)
        start = torch.full((B,), P, dtype=torch.int32, device="cuda")
        out = eng.forward_uniform(idc[:, P:].contiguous(), start, slots, P + q, last_only=last)
        torch.cuda.synchronize()
        assert eng.tp_error() == 0
        outs.append(out.cpu())
        eng.close()
    ref = qwen2_forward({k: v.cpu() for k, v in w.items()}, cfg, ids, last_n=1 if last else q)
    ref = ref.reshape(outs[1].shape)
    assert (outs[1] - ref).abs().max().item() <= 2e-2
    assert (outs[1] - outs[0]).abs().max().item() <= 2e-2
