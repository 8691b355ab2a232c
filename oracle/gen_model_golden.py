"""Pin oracle/model_oracle.py to the third-party implementation the reference relies on.

Builds a tiny random Qwen2ForCausalLM with the installed ``transformers`` (the reference pins
transformers>=4.40,<5, requirements.txt:5-6; this container has a newer one - the version is recorded
in the fixture), runs it in fp32 on CPU and stores weights (bf16 bit patterns), token ids and logits in
tests/golden/qwen2_tiny_golden.npz.  Run in the build container: ``python oracle/gen_model_golden.py``.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import transformers
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(0)
    cfg = Qwen2Config(hidden_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      intermediate_size=256, vocab_size=512, max_position_embeddings=512, rms_norm_eps=1e-6,
                      rope_theta=1e6, tie_word_embeddings=False, attn_implementation="eager")
    cfg.head_dim = 64
    try:
        cfg.rope_parameters = {"rope_type": "default", "rope_theta": 1e6}
    except Exception:
        pass
    model = Qwen2ForCausalLM(cfg).eval()
    sd = {}
    for k, v in model.state_dict().items():
        v = (v.float() * (1.0 if "norm" in k else 1.0))
        if "bias" in k:
            v = torch.randn_like(v) * 0.02            # HF zero-inits biases: make them matter
        sd[k] = v.to(torch.bfloat16)
    model.load_state_dict({k: v.float() for k, v in sd.items()})
    ids = torch.randint(0, cfg.vocab_size, (3, 37), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        logits = model(ids).logits.float()
    out = {"input_ids": ids.numpy().astype(np.int32), "logits": logits.numpy(),
           "transformers_version": np.array(transformers.__version__),
           "torch_version": np.array(torch.__version__)}
    for k, v in sd.items():
        out["w::" + k] = v.view(torch.int16).numpy()
    path = os.path.join(ROOT, "tests", "golden", "qwen2_tiny_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; transformers", transformers.__version__)


if __name__ == "__main__":
    main()
