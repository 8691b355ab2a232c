"""oracle/model_oracle.py pinned to HF Qwen2ForCausalLM (fixture made by oracle/gen_model_golden.py)."""
import os

import numpy as np
import pytest
import torch

from asd_b200.models.qwen2 import tiny_config
from oracle.model_oracle import qwen2_forward

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qwen2_tiny_golden.npz")


def load_golden():
    z = np.load(GOLD)
    w = {k[3:]: torch.from_numpy(z[k]).view(torch.bfloat16) for k in z.files if k.startswith("w::")}
    return w, torch.from_numpy(z["input_ids"]).long(), torch.from_numpy(z["logits"])


def test_oracle_matches_hf_golden():
    w, ids, logits = load_golden()
    got = qwen2_forward(w, tiny_config(), ids)
    assert got.shape == logits.shape
    assert (got - logits).abs().max().item() < 2e-4       # fp32 vs fp32: summation order only
    assert (got.argmax(-1) == logits.argmax(-1)).float().mean().item() == 1.0


def test_oracle_matches_live_hf():
    tr = pytest.importorskip("transformers")
    cfg = tr.Qwen2Config(hidden_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                         intermediate_size=256, vocab_size=512, max_position_embeddings=512, rms_norm_eps=1e-6,
                         rope_theta=1e6, tie_word_embeddings=True, attn_implementation="eager")
    cfg.head_dim = 64
    torch.manual_seed(3)
    model = tr.Qwen2ForCausalLM(cfg).eval()
    sd = {k: v.float() for k, v in model.state_dict().items()}
    ids = torch.randint(0, 512, (2, 19))
    with torch.no_grad():
        ref = model(ids).logits.float()
    got = qwen2_forward(sd, tiny_config(tie_word_embeddings=True), ids)
    assert (got - ref).abs().max().item() < 2e-4
