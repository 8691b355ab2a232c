"""Host side of the draft-then-verify engine: device-memory ownership (weights, paged KV pool, page
table) and the step loop.  All arithmetic is in libasd_b200.so; torch is used for tensors, streams
and small index bookkeeping on the device (no host synchronisation inside a step)."""
from __future__ import annotations

import ctypes
import math
import threading
from typing import Dict, Optional

import torch

from ._lib import AsdError, ModelConfigC, check, lib
from .models.qwen2 import Qwen2Config, pack_layer, random_packed_layer
from .ops import NUM_FEATURES


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class QwenEngine:
    """One model on one tensor-parallel rank: wraps an ``asd_engine_t`` handle."""

    def __init__(self, cfg: Qwen2Config, max_seqs: int, max_seq_len: int, max_tokens: int = 256, page_size: int = 16,
                 tp_rank: int = 0, tp_size: int = 1, device="cuda", shuffle_pages: bool = True,
                 fuse_norm: bool = True):
        if not torch.cuda.is_available():
            raise AsdError("QwenEngine needs a CUDA device (there is no CPU fallback)")
        self.cfg, self.device = cfg, torch.device(device)
        self.tp_rank, self.tp_size = tp_rank, tp_size
        self.max_seqs, self.max_seq_len, self.max_tokens, self.page_size = max_seqs, max_seq_len, max_tokens, page_size
        assert cfg.num_attention_heads % tp_size == 0 and cfg.num_key_value_heads % tp_size == 0
        self.c = ModelConfigC(cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads // tp_size,
                              cfg.num_key_value_heads // tp_size, cfg.head_dim, cfg.intermediate_size // tp_size,
                              cfg.vocab_size, cfg.rms_norm_eps, max_tokens, page_size, tp_rank, tp_size)
        self._lock = threading.Lock()
        with torch.cuda.device(self.device):
            self.h = lib().asd_engine_create(ctypes.byref(self.c))
        if not self.h:
            check(-1, "asd_engine_create")
        self.max_pages = (max_seq_len + page_size - 1) // page_size
        self.num_pages = self.max_pages * max_seqs
        nbytes = lib().asd_engine_kv_pool_bytes(ctypes.byref(self.c), self.num_pages)
        self.kv_pool = torch.zeros(nbytes // 2, dtype=torch.bfloat16, device=self.device)
        ids = torch.arange(self.num_pages, dtype=torch.int32)
        if shuffle_pages:  # pages of a sequence are deliberately NOT contiguous: the indirection is always live
            ids = ids[torch.randperm(self.num_pages, generator=torch.Generator().manual_seed(1234))]
        self.page_table = ids.view(max_seqs, self.max_pages).contiguous().to(self.device)
        check(lib().asd_engine_set_kv(self.h, self.kv_pool.data_ptr(), self.num_pages, self.page_table.data_ptr(),
                                      max_seqs, self.max_pages), "asd_engine_set_kv")
        inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.int64).float() / cfg.head_dim))
        self.inv_freq = inv.to(self.device)
        self._weights = []   # keep tensors alive
        self._index_cache = {}
        self.fuse_norm = bool(fuse_norm)   # RMSNorm fused into the GEMM epilogues (no separate norm kernels)
        self.set_option("fuse_norm", int(self.fuse_norm))
        self._globals_set = False

    def close(self):
        if getattr(self, "h", None):
            lib().asd_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def _set_layer(self, l: int, t: Dict[str, torch.Tensor]):
        self._weights.append(t)
        check(lib().asd_engine_set_layer(self.h, l, t["wqkv"].data_ptr(), t["bqkv"].data_ptr(), t["wo"].data_ptr(),
                                         t["wgateup"].data_ptr(), t["wdown"].data_ptr(), t["ln1"].data_ptr(),
                                         t["ln2"].data_ptr()), "asd_engine_set_layer")

    def _set_globals(self, embed, final_norm, lm_head):
        self._weights.append((embed, final_norm, lm_head))
        check(lib().asd_engine_set_globals(self.h, embed.data_ptr(), final_norm.data_ptr(), lm_head.data_ptr(),
                                           self.inv_freq.data_ptr()), "asd_engine_set_globals")

    def load_hf_weights(self, w: Dict[str, torch.Tensor]):
        """HF-named Qwen2 state dict (any device) -> packed bf16 engine layout on this rank."""
        for l in range(self.cfg.num_hidden_layers):
            self._set_layer(l, pack_layer(w, self.cfg, l, self.tp_rank, self.tp_size, self.device))
        dev = lambda t: t.to(device=self.device, dtype=torch.bfloat16).contiguous()
        embed = dev(w["model.embed_tokens.weight"])
        head = embed if "lm_head.weight" not in w else dev(w["lm_head.weight"])
        self._set_globals(embed, dev(w["model.norm.weight"]), head)
        return self

    def load_random(self, seed: int, std: float = 0.02):
        """Random-init weights generated directly in the engine layout (full-size benchmarks)."""
        g = torch.Generator(device=self.device).manual_seed(seed * 1000 + self.tp_rank)
        for l in range(self.cfg.num_hidden_layers):
            self._set_layer(l, random_packed_layer(self.cfg, g, self.tp_size, self.device, std))
        gg = torch.Generator(device=self.device).manual_seed(seed * 1000 + 999)   # replicated tensors: same on all ranks
        rn = lambda *s, mean=0.0: (torch.randn(*s, generator=gg, device=self.device) * std + mean).to(torch.bfloat16)
        embed = rn(self.cfg.vocab_size, self.cfg.hidden_size)
        head = embed if self.cfg.tie_word_embeddings else rn(self.cfg.vocab_size, self.cfg.hidden_size)
        self._set_globals(embed, rn(self.cfg.hidden_size, mean=1.0), head)
        return self

    def set_option(self, name: str, value: int):
        check(lib().asd_engine_set_option(self.h, name.encode(), int(value)), "asd_engine_set_option")

    PROFILE_CLASSES = ("gemm", "attention", "glue", "allreduce", "lm_head")

    def profile_read(self):
        """(ms, launches) per kernel class since the last read; needs set_option('profile', 1)."""
        ms = (ctypes.c_float * 5)()
        n = (ctypes.c_int * 5)()
        check(lib().asd_engine_profile_read(self.h, ms, n, 5), "asd_engine_profile_read")
        return {c: (ms[i], n[i]) for i, c in enumerate(self.PROFILE_CLASSES)}

    def enable_p2p(self, group=None):
        """Exchange CUDA-IPC handles of the TP receive buffers with the other ranks (torch.distributed) and
        switch the row-parallel boundaries to the peer-memory paths: the all-reduce inside the O / down GEMM
        epilogues where it pays (2 ranks, small exchanges), else the one-kernel all-reduce + residual + norm
        (options ``tp_fused`` / ``tp_two_shot``)."""
        import torch.distributed as dist
        mine = ctypes.create_string_buffer(3 * 64)
        check(lib().asd_engine_ipc_export(self.h, mine), "asd_engine_ipc_export")
        table = [None] * self.tp_size
        dist.all_gather_object(table, bytes(mine.raw), group=group)
        blob = ctypes.create_string_buffer(b"".join(table), 3 * 64 * self.tp_size)
        with torch.cuda.device(self.device):
            check(lib().asd_engine_ipc_import(self.h, blob), "asd_engine_ipc_import")
        dist.barrier(group=group)
        return self

    def tp_error(self) -> int:
        return int(lib().asd_engine_tp_error(self.h))

    def set_allreduce(self, comm_ptr: int, fn_ptr: int):
        check(lib().asd_engine_set_allreduce(self.h, ctypes.c_void_p(comm_ptr), ctypes.c_void_p(fn_ptr)),
              "asd_engine_set_allreduce")

    # ------------------------------------------------------------------ forward
    def forward(self, tokens, positions, token_slot, cu_q, seq_slot, max_qlen: int, max_kv_len: int,
                logit_rows: Optional[torch.Tensor] = None, logits_out: Optional[torch.Tensor] = None,
                want_logits: bool = True, logits_ld: int = 0):
        """int32 CUDA tensors; returns fp32 logits [rows, vocab] (or None)."""
        M, nseq = tokens.numel(), seq_slot.numel()
        for t in (tokens, positions, token_slot, cu_q, seq_slot):
            assert t.dtype == torch.int32 and t.is_contiguous() and t.is_cuda, "int32 contiguous CUDA tensors"
        n_rows = 0 if not want_logits else (M if logit_rows is None else logit_rows.numel())
        if n_rows and logits_out is None:
            logits_out = torch.empty(n_rows, self.cfg.vocab_size, dtype=torch.float32, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            rc = lib().asd_engine_forward(
                self.h, tokens.data_ptr(), positions.data_ptr(), token_slot.data_ptr(), M, cu_q.data_ptr(),
                seq_slot.data_ptr(), nseq, max_qlen, max_kv_len,
                None if logit_rows is None else logit_rows.data_ptr(), n_rows,
                None if not n_rows else logits_out.data_ptr(), logits_ld, _stream())
        check(rc, "asd_engine_forward")
        return logits_out if n_rows else None

    def _uniform_index(self, nseq: int, q: int):
        """static index tensors of a uniform (nseq x q) call, built once per shape"""
        key = (nseq, q)
        c = self._index_cache.get(key)
        if c is None:
            ar = torch.arange(q, dtype=torch.int32, device=self.device)
            c = dict(ar=ar[None].contiguous(),
                     cu_q=torch.arange(0, (nseq + 1) * q, q, dtype=torch.int32, device=self.device),
                     last_rows=torch.arange(q - 1, nseq * q, q, dtype=torch.int32, device=self.device))
            self._index_cache[key] = c
        return c

    def forward_uniform(self, tokens2d: torch.Tensor, start_pos: torch.Tensor, slots: torch.Tensor, max_kv_len: int,
                        last_only: bool = False, want_logits: bool = True, logits_out=None, logits_ld: int = 0):
        """tokens2d int32 [nseq, q]; start_pos int32 [nseq] (position of column 0); slots int32 [nseq]."""
        nseq, q = tokens2d.shape
        ix = self._uniform_index(nseq, q)
        positions = (start_pos[:, None] + ix["ar"]).reshape(-1)
        token_slot = slots[:, None].expand(nseq, q).reshape(-1) if q > 1 else slots
        rows = ix["last_rows"] if (last_only and want_logits and q > 1) else None
        return self.forward(tokens2d.reshape(-1).contiguous(), positions.contiguous(), token_slot.contiguous(),
                            ix["cu_q"], slots.contiguous(), q,
                            max_kv_len, rows, logits_out, want_logits, logits_ld)

    def prefill(self, prompt_ids: torch.Tensor, slots: torch.Tensor, chunk: int = 0, want_logits: bool = True):
        """Chunked prefill of equal-length prompts [nseq, P]; returns the logits of the last position."""
        nseq, P = prompt_ids.shape
        G = self.cfg.num_attention_heads // self.cfg.num_key_value_heads
        if chunk <= 0:
            chunk = max(1, min(128 // G, self.max_tokens // nseq))
        logits = None
        zero = torch.zeros(nseq, dtype=torch.int32, device=self.device)
        for s in range(0, P, chunk):
            e = min(P, s + chunk)
            last = e == P
            logits = self.forward_uniform(prompt_ids[:, s:e].to(torch.int32), zero + s, slots, e, last_only=True,
                                          want_logits=want_logits and last)
        return logits


class SpecDecoder:
    """Chain draft-then-verify: k draft steps on the draft engine, ONE (k+1)-token verify forward on
    the target engine, ONE fused rejection-sampling launch; state stays on the device.

    ``draft=None`` degenerates to plain autoregressive decoding of the target (k = 0)."""

    def __init__(self, target: QwenEngine, draft: Optional[QwenEngine], batch: int, k: int, temperature: float,
                 seed: int = 4321):
        self.t, self.d = target, draft
        self.B, self.k, self.T = batch, (k if draft is not None else 0), float(temperature)
        dev = target.device
        self.device = dev
        V = target.cfg.vocab_size
        if draft is not None:
            assert draft.cfg.vocab_size == V, "draft and target must share the vocabulary"
        from .ops import RejectionSampler
        self.sampler = RejectionSampler(batch, self.k, dev)
        self.row_sampler = RejectionSampler(batch, 0, dev)
        self.slots = torch.arange(batch, dtype=torch.int32, device=dev)
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.target_logits = torch.empty(batch, self.k + 1, V, dtype=torch.float32, device=dev)
        self.draft_logits = torch.empty(batch, max(self.k, 1), V, dtype=torch.float32, device=dev)
        self.draft_tokens = torch.zeros(batch, max(self.k, 1), dtype=torch.int32, device=dev)
        self.pos = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.last_tok = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.prev_tok = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.kv_bound = 0     # host-side upper bound of every sequence length (no sync needed)
        self.time_verify = False
        self.verify_events = []
        self._empty_i = torch.zeros(batch, 0, dtype=torch.int32, device=dev)
        self._empty_d = torch.zeros(batch, 0, dtype=torch.float64, device=dev)

    def _uniform(self, *shape):
        return torch.rand(*shape, dtype=torch.float64, device=self.device, generator=self.gen)

    def _sample_rows(self, logits_bv: torch.Tensor) -> torch.Tensor:
        """one token per row from softmax(logits/T) (argmax when T <= 0), fused kernel with k = 0"""
        out = self.row_sampler(logits_bv, None, self._empty_i, self._empty_d, self._uniform(self.B), self.T)
        return out["out_tokens"][:, 0].clone()

    def prefill(self, prompt_ids: torch.Tensor):
        """prompt_ids int [B, P] (equal lengths).  Fills both KV caches and samples the first token."""
        B, P = prompt_ids.shape
        assert B == self.B and P >= 1
        prompt_ids = prompt_ids.to(self.device)
        logits = self.t.prefill(prompt_ids, self.slots)
        if self.d is not None:
            self.d.prefill(prompt_ids, self.slots, want_logits=False)
        self.last_tok = self._sample_rows(logits.contiguous())
        self.prev_tok = prompt_ids[:, -1].to(torch.int32).contiguous()
        self.pos = torch.full((B,), P, dtype=torch.int32, device=self.device)
        self.kv_bound = P + 1
        return self.last_tok

    def seed_state(self, prefix_len: int, last_tok: torch.Tensor, prev_tok: torch.Tensor):
        """Benchmark helper: declare that both KV caches already hold ``prefix_len`` positions."""
        self.pos = torch.full((self.B,), prefix_len, dtype=torch.int32, device=self.device)
        self.last_tok, self.prev_tok = last_tok.to(torch.int32).contiguous(), prev_tok.to(torch.int32).contiguous()
        self.kv_bound = prefix_len + 1

    def step(self):
        """One draft-then-verify step for the whole batch.  Returns the sampler's output dict
        (device tensors): out_tokens [B, k+1] (-1 padded), accepted_len [B], accept_mask, features."""
        B, k, V = self.B, self.k, self.t.cfg.vocab_size
        bound = self.kv_bound + k + 1
        if k > 0:
            # draft step 1 re-feeds the previous token so the draft KV is complete after an all-accept
            two = torch.stack([self.prev_tok, self.last_tok], 1)
            self.d.forward_uniform(two, self.pos - 1, self.slots, bound, last_only=True,
                                   logits_out=self.draft_logits, logits_ld=k * V)
            x = self._sample_rows_strided(0)
            self.draft_tokens[:, 0] = x
            for i in range(1, k):
                self.d.forward_uniform(x[:, None], self.pos + i, self.slots, bound,
                                       logits_out=self.draft_logits[:, i], logits_ld=k * V)
                x = self._sample_rows_strided(i)
                self.draft_tokens[:, i] = x
            toks = torch.cat([self.last_tok[:, None], self.draft_tokens], 1)
        else:
            toks = self.last_tok[:, None]
        if self.time_verify:     # two CUDA events per step on the launching stream (bench.py)
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        self.t.forward_uniform(toks, self.pos, self.slots, bound, logits_out=self.target_logits)
        if self.time_verify:
            ev[1].record()
            self.verify_events.append(ev)
        out = self.sampler(self.target_logits, self.draft_logits if k > 0 and self.T > 0 else None,
                           self.draft_tokens[:, :k].contiguous() if k > 0 else self._empty_i,
                           self._uniform(B, k) if k > 0 else self._empty_d, self._uniform(B), self.T)
        n = out["accepted_len"].to(torch.int64)
        new_last = out["out_tokens"].gather(1, n[:, None])[:, 0]
        self.prev_tok = toks.gather(1, n[:, None])[:, 0].contiguous()
        self.last_tok = new_last.contiguous()
        self.pos = self.pos + out["accepted_len"] + 1
        self.kv_bound = bound
        return out

    def step_host(self, host_state: torch.Tensor, host_tokens: torch.Tensor, host_accepted: torch.Tensor):
        """Same step driven from HOST (pinned) buffers, the way a non-torch integrator would call it:
        host_state int32 [3, B] = (last_tok, prev_tok, pos) is copied to the device, the step runs, then
        the emitted tokens [B, k+1], accepted lengths [B] and the next state (in place) are copied back;
        returns after the copy-out has completed."""
        dev_state = host_state.to(self.device, non_blocking=True)
        self.last_tok, self.prev_tok, self.pos = (dev_state[0].contiguous(), dev_state[1].contiguous(),
                                                  dev_state[2].contiguous())
        out = self.step()
        host_tokens.copy_(out["out_tokens"], non_blocking=True)
        host_accepted.copy_(out["accepted_len"], non_blocking=True)
        host_state.copy_(torch.stack([self.last_tok, self.prev_tok, self.pos]), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_tokens, host_accepted

    def _sample_rows_strided(self, i: int) -> torch.Tensor:
        # the fused sampler wants contiguous rows: draft step i wrote rows with stride k*V
        return self._sample_rows(self.draft_logits[:, i].contiguous())
