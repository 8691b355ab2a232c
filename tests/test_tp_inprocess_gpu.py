"""In-process tensor parallelism behind the reference-facing seam: ``TPQwenEngine`` and
``Stage(tensor_parallel_size=t, gpu_ids=[...])`` - t rank engines on t GPUs of ONE process, row-parallel
boundaries over peer-mapped memory - against the CPU oracle, against the single-GPU engine at the BASELINE
widths, and through ``StageManager`` + ``AdaptiveSpeculativePipeline`` with the reference's placement
(7B on its own GPU, 32B TP=2, 72B TP=4; configs/qwen3_models.yaml:10-51).  Skipped below the GPU count needed."""
from dataclasses import replace

import pytest
import torch

from asd_b200.models.qwen2 import QWEN25, Qwen2Config, random_hf_weights

pytestmark = pytest.mark.gpu
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def _decidable_agree(got, ref, tol=2e-2):
    top2 = ref.topk(2, -1).values
    dec = (top2[..., 0] - top2[..., 1]) > 2 * tol
    return float((got.argmax(-1) == ref.argmax(-1))[dec].float().mean()) if dec.any() else 1.0


@pytest.mark.skipif(NGPU < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["fused", "kernel", "two_shot"])
def test_tp2_inprocess_matches_oracle(mode):
    from asd_b200.engine import TPQwenEngine
    from oracle.model_oracle import qwen2_forward
    cfg = Qwen2Config(512, 2, 8, 2, 1024, 4096, head_dim=128, name="tp-test")
    w = random_hf_weights(cfg, seed=11, logit_std=0.4)
    ids = torch.randint(0, cfg.vocab_size, (3, 70), generator=torch.Generator().manual_seed(5))
    eng = TPQwenEngine(cfg, [0, 1], max_seqs=3, max_seq_len=96, max_tokens=64).load_hf_weights(w)
    if mode != "fused":
        eng.set_option("tp_fused", 0)
        eng.set_option("tp_two_shot", 1 if mode == "two_shot" else -1)
    slots = torch.arange(3, dtype=torch.int32, device="cuda:0")
    idc = ids.to("cuda:0").to(torch.int32)
    eng.prefill(idc[:, :64], slots)
    ver = eng.forward_uniform(idc[:, 64:].contiguous(), torch.full((3,), 64, dtype=torch.int32, device="cuda:0"), slots, 70)
    for d in range(2):
        torch.cuda.synchronize(d)
    got = ver.view(3, 6, -1).cpu()
    ref = qwen2_forward(w, cfg, ids)[:, 64:]
    assert (got - ref).abs().max().item() <= 2e-2
    assert _decidable_agree(got, ref) >= 0.999
    assert eng.tp_error() == 0
    eng.close()


@pytest.mark.skipif(NGPU < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("tp", [2, 4, 8])
def test_tp_inprocess_32b_width_matches_single_gpu(tp):
    """2 layers of the 32B shape at M = 96 (the verify shape of BASELINE configs[2]): the sharded engine must
    agree with the single-GPU engine (itself checked against the oracle in test_engine_baseline_shapes_gpu.py)
    within the north-star tolerance.  The two differ only in the fp32 summation order of the row-parallel partials,
    but an fp32 difference in the last bit flips bf16 roundings of the next GEMM's operand, so the worst of the
    14.6 million compared logits sits at about half the bf16-vs-fp32 error (measured 0.012 at max|logit| 2.1)."""
    if NGPU < tp:
        pytest.skip(f"needs {tp} GPUs")
    from asd_b200.engine import QwenEngine, TPQwenEngine
    cfg = replace(QWEN25["32b"], num_hidden_layers=2)
    w = random_hf_weights(cfg, seed=3, device="cuda:0", logit_std=0.25)
    B, P, q = 16, 64, 6
    ids = torch.randint(0, cfg.vocab_size, (B, P + q), generator=torch.Generator().manual_seed(7)).to("cuda:0").to(torch.int32)
    slots = torch.arange(B, dtype=torch.int32, device="cuda:0")
    start = torch.full((B,), P, dtype=torch.int32, device="cuda:0")
    outs = []
    for make in (lambda: QwenEngine(cfg, max_seqs=B, max_seq_len=P + q + 16, max_tokens=256, device="cuda:0"),
                 lambda: TPQwenEngine(cfg, list(range(tp)), max_seqs=B, max_seq_len=P + q + 16, max_tokens=256)):
        eng = make().load_hf_weights(w)
        eng.prefill(ids[:, :P], slots, want_logits=False)
        ver = eng.forward_uniform(ids[:, P:].contiguous(), start, slots, P + q)
        for d in range(NGPU):
            torch.cuda.synchronize(d)
        outs.append(ver.view(B, q, -1).cpu())
        if hasattr(eng, "tp_error"):
            assert eng.tp_error() == 0
        eng.close()
    err = (outs[0] - outs[1]).abs().max().item()
    assert err <= 2e-2, err
    assert _decidable_agree(outs[1], outs[0]) >= 0.999


@pytest.mark.skipif(NGPU < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("tp,two_shot", [(2, 0), (2, 1), (4, 0), (4, -1), (8, 0)])
def test_tp_inprocess_large_verify_matches_single_gpu(tp, two_shot):
    """M = 320 tokens (> 256: the tensor-bound CTA-pair GEMM, BASELINE configs[4]'s regime; 6.5 MB per boundary): O
    through the weight-streaming kernel with the fused exchange (two ranks) or the all-reduce kernel, down through the
    CTA-pair kernel + slice reduction + all-reduce kernel (one-shot / two-shot, tp_two_shot forces either).  All must
    agree with one GPU."""
    if NGPU < tp:
        pytest.skip(f"needs {tp} GPUs")
    from asd_b200.engine import QwenEngine, TPQwenEngine
    cfg = replace(QWEN25["32b"], num_hidden_layers=2)
    w = random_hf_weights(cfg, seed=3, device="cuda:0", logit_std=0.25)
    B, P, q = 32, 48, 10
    ids = torch.randint(0, cfg.vocab_size, (B, P + q), generator=torch.Generator().manual_seed(9)).to("cuda:0").to(torch.int32)
    slots = torch.arange(B, dtype=torch.int32, device="cuda:0")
    start = torch.full((B,), P, dtype=torch.int32, device="cuda:0")
    outs = []
    for make in (lambda: QwenEngine(cfg, max_seqs=B, max_seq_len=P + q + 16, max_tokens=512, device="cuda:0"),
                 lambda: TPQwenEngine(cfg, list(range(tp)), max_seqs=B, max_seq_len=P + q + 16, max_tokens=512)):
        eng = make().load_hf_weights(w)
        if hasattr(eng, "tp_error"):
            eng.set_option("tp_two_shot", two_shot)
        eng.prefill(ids[:, :P], slots, want_logits=False)
        for _ in range(3):      # several forwards: the receive areas alternate with the epoch parity
            ver = eng.forward_uniform(ids[:, P:].contiguous(), start, slots, P + q)
        for d in range(NGPU):
            torch.cuda.synchronize(d)
        outs.append(ver.view(B, q, -1).cpu())
        if hasattr(eng, "tp_error"):
            assert eng.tp_error() == 0
        eng.close()
    err = (outs[0] - outs[1]).abs().max().item()
    assert err <= 2e-2, err
    assert _decidable_agree(outs[1], outs[0]) >= 0.999


def _cascade_placement():
    """the reference's placement on as many GPUs as the box has (every stage on its own GPUs)"""
    if NGPU >= 7:
        return [("7b", 1, [0]), ("32b", 2, [1, 2]), ("72b", 4, [3, 4, 5, 6])]
    if NGPU >= 3:
        return [("7b", 1, [0]), ("32b", 2, [1, 2])]
    return None


@pytest.mark.skipif(NGPU < 3, reason="needs >= 3 GPUs (7B on GPU 0, 32B TP=2 on GPUs 1-2; 72B TP=4 with >= 7)")
def test_stage_manager_cascade_with_tensor_parallel_stages():
    """BASELINE configs[3]: 7B -> 32B (TP=2) -> 72B (TP=4) on disjoint GPUs through StageManager and the pipeline,
    layer-truncated (2 layers each) at the real widths and vocabulary."""
    from asd_b200.models.predictor import FeatureExtractor, QualityPredictor
    from asd_b200.models.stage import StageConfig, StageManager
    from asd_b200.serving.pipeline import AdaptiveSpeculativePipeline, PipelineConfig
    place = _cascade_placement()
    scs = [StageConfig(f"qwen2.5-{s}", s, tp, gpu_ids=g, config=replace(QWEN25[s], num_hidden_layers=2))
           for s, tp, g in place]
    mgr = StageManager(scs, k=3, stage_kwargs=dict(max_batch=4, max_model_len=256))
    info = mgr.get_stage(place[-1][0]).get_model_info()
    assert info["tensor_parallel_size"] == place[-1][1] and info["gpu_ids"] == place[-1][2]
    pipe = AdaptiveSpeculativePipeline(mgr, QualityPredictor(256), FeatureExtractor(),
                                       PipelineConfig(lambda_value=1e6, risk_adjustment=True))
    r = pipe.process_request("What is the capital of France?", max_tokens=12, temperature=0.7)
    assert 0 <= r.stopped_at_stage < len(place) and isinstance(r.output, str)
    assert len(r.stage_probabilities) == r.stopped_at_stage + 1
    # greedy through the sharded target with a draft on another GPU == greedy of the same target alone
    last = mgr.get_stage(place[-1][0])
    t_spec, _, st_spec = last.generate(["hello world", "a longer prompt, different length"], max_tokens=10, temperature=0.0)
    saved, last.draft = last.draft, None
    t_plain, _, _ = last.generate(["hello world", "a longer prompt, different length"], max_tokens=10, temperature=0.0)
    last.draft = saved
    assert t_spec == t_plain and st_spec["decode_steps"] >= 1
    res = pipe.batch_process(["one", "two two", "three three three"], max_tokens=8, batched=True)
    assert len(res) == 3
    pipe.shutdown()
