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


class _BatchOps:
    """Uniform-batch / prefill helpers shared by ``QwenEngine`` and ``TPQwenEngine`` (anything with ``forward``,
    ``device``, ``cfg``, ``max_tokens`` and an ``_index_cache`` dict)."""

    def _uniform_index(self, nseq: int, q: int):
        """static index tensors of a uniform (nseq x q) call, built once per shape"""
        key = (nseq, q)
        c = self._index_cache.get(key)
        if c is None:
            c = self._index_cache[key] = {}
        return c

    def _uniform_index_on(self, nseq: int, q: int, dev):
        """index tensors live on the CALLER's device (the engine hops them over if it sits elsewhere)"""
        c = self._uniform_index(nseq, q)
        d = c.get(dev)
        if d is None:
            ar = torch.arange(q, dtype=torch.int32, device=dev)
            d = c[dev] = dict(ar=ar[None].contiguous(),
                              cu_q=torch.arange(0, (nseq + 1) * q, q, dtype=torch.int32, device=dev),
                              last_rows=torch.arange(q - 1, nseq * q, q, dtype=torch.int32, device=dev))
        return d

    def forward_uniform(self, tokens2d: torch.Tensor, start_pos: torch.Tensor, slots: torch.Tensor, max_kv_len: int,
                        last_only: bool = False, want_logits: bool = True, logits_out=None, logits_ld: int = 0):
        """tokens2d int32 [nseq, q]; start_pos int32 [nseq] (position of column 0); slots int32 [nseq]."""
        nseq, q = tokens2d.shape
        ix = self._uniform_index_on(nseq, q, tokens2d.device)
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
        zero = torch.zeros(nseq, dtype=torch.int32, device=prompt_ids.device)
        for s in range(0, P, chunk):
            e = min(P, s + chunk)
            last = e == P
            logits = self.forward_uniform(prompt_ids[:, s:e].to(torch.int32), zero + s, slots, e, last_only=True,
                                          want_logits=want_logits and last)
        return logits

    def prefill_ragged(self, prompts, slots: torch.Tensor, chunk: int = 0, device=None):
        """Chunked prefill of prompts of DIFFERENT lengths in one batch (the reference's vLLM stage takes ragged
        prompt lists, docs/guides/RESEARCH_PROTOCOL.md:272-284).  ``prompts``: list of token-id lists; returns the
        logits of each sequence's last position [nseq, V].  Every chunk is one ragged forward (cu_q / per-token
        positions); a sequence drops out of the batch once its prompt is consumed."""
        dev = torch.device(device) if device is not None else slots.device
        nseq = len(prompts)
        lens = [len(p) for p in prompts]
        assert nseq == slots.numel() and min(lens) >= 1
        G = self.cfg.num_attention_heads // self.cfg.num_key_value_heads
        if chunk <= 0:
            chunk = max(1, min(128 // G, self.max_tokens // nseq))
        out = torch.empty(nseq, self.cfg.vocab_size, dtype=torch.float32, device=dev)
        slots_h = slots.cpu().tolist()
        for s0 in range(0, max(lens), chunk):
            toks, pos, tslot, cu, sslot, rows, owners = [], [], [], [0], [], [], []
            for b in range(nseq):
                e = min(lens[b], s0 + chunk)
                if e <= s0:
                    continue
                toks += prompts[b][s0:e]
                pos += list(range(s0, e))
                tslot += [slots_h[b]] * (e - s0)
                cu.append(len(toks))
                sslot.append(slots_h[b])
                if e == lens[b]:
                    rows.append(len(toks) - 1)
                    owners.append(b)
            mk = lambda x: torch.tensor(x, dtype=torch.int32).to(dev, non_blocking=True)
            lg = self.forward(mk(toks), mk(pos), mk(tslot), mk(cu), mk(sslot), min(chunk, max(lens) - s0),
                              min(max(lens), s0 + chunk), mk(rows) if rows else None, None, bool(rows))
            if rows:
                out[torch.tensor(owners, device=dev)] = lg
        return out


class QwenEngine(_BatchOps):
    """One model on one tensor-parallel rank: wraps an ``asd_engine_t`` handle."""

    def __init__(self, cfg: Qwen2Config, max_seqs: int, max_seq_len: int, max_tokens: int = 256, page_size: int = 16,
                 tp_rank: int = 0, tp_size: int = 1, device="cuda", shuffle_pages: bool = True,
                 fuse_norm: bool = True):
        if not torch.cuda.is_available():
            raise AsdError("QwenEngine needs a CUDA device (there is no CPU fallback)")
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.cfg, self.device = cfg, dev
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
        gg = torch.Generator(device=self.device).manual_seed(seed * 1000 + 999)   # replicated tensors: same on all ranks
        for l in range(self.cfg.num_hidden_layers):
            self._set_layer(l, random_packed_layer(self.cfg, g, self.tp_size, self.device, std, gg))
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
        """int32 CUDA tensors; returns fp32 logits [rows, vocab] (or None).  Inputs that live on another GPU (a
        draft stage placed on its own device, configs/qwen3_models.yaml:15,39,51) are copied over NVLink on a
        side stream of this engine's device and the logits are copied back; both streams are ordered by events,
        the host never waits."""
        M, nseq = tokens.numel(), seq_slot.numel()
        for t in (tokens, positions, token_slot, cu_q, seq_slot):
            assert t.dtype == torch.int32 and t.is_contiguous() and t.is_cuda, "int32 contiguous CUDA tensors"
        if max_kv_len > self.max_pages * self.page_size:
            raise AsdError(f"sequence length bound {max_kv_len} exceeds this engine's max_seq_len "
                           f"{self.max_pages * self.page_size} (page-table overrun refused)")
        n_rows = 0 if not want_logits else (M if logit_rows is None else logit_rows.numel())
        if tokens.device != self.device:
            return _remote_forward(self, self.device, tokens, positions, token_slot, cu_q, seq_slot, max_qlen,
                                   max_kv_len, logit_rows, logits_out, want_logits, logits_ld)
        if n_rows and logits_out is None:
            logits_out = torch.empty(n_rows, self.cfg.vocab_size, dtype=torch.float32, device=self.device)
        if n_rows and logits_out.device != self.device:
            raise AsdError("logits_out must live on the engine's device")
        with self._lock, torch.cuda.device(self.device):
            rc = lib().asd_engine_forward(
                self.h, tokens.data_ptr(), positions.data_ptr(), token_slot.data_ptr(), M, cu_q.data_ptr(),
                seq_slot.data_ptr(), nseq, max_qlen, max_kv_len,
                None if logit_rows is None else logit_rows.data_ptr(), n_rows,
                None if not n_rows else logits_out.data_ptr(), logits_ld, _stream())
        check(rc, "asd_engine_forward")
        return logits_out if n_rows else None

    def _hop_stream(self):
        """side stream of this engine's device for callers whose current stream is on another GPU"""
        st = getattr(self, "_side_stream", None)
        if st is None:
            st = self._side_stream = torch.cuda.Stream(device=self.device)
        return st

    def _hop_logits(self, n_rows: int):
        buf = getattr(self, "_side_logits", None)
        if buf is None or buf.shape[0] < n_rows:
            buf = self._side_logits = torch.empty(max(n_rows, 1), self.cfg.vocab_size, dtype=torch.float32,
                                                  device=self.device)
        return buf[:n_rows]


def _remote_forward(eng, dev, tokens, positions, token_slot, cu_q, seq_slot, max_qlen, max_kv_len, logit_rows,
                    logits_out, want_logits, logits_ld):
    """Run ``eng.forward`` on ``dev`` for a caller whose tensors (and current stream) are on another GPU."""
    src = tokens.device
    n_rows = 0 if not want_logits else (tokens.numel() if logit_rows is None else logit_rows.numel())
    V = eng.cfg.vocab_size
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(src))
    side = eng._hop_stream()
    ins = (tokens, positions, token_slot, cu_q, seq_slot, logit_rows)
    with torch.cuda.device(dev), torch.cuda.stream(side):
        side.wait_event(ready)
        loc = []
        for t in ins:
            if t is None:
                loc.append(None)
                continue
            loc.append(t.to(dev, non_blocking=True))
            t.record_stream(side)
        local = eng._hop_logits(n_rows) if n_rows else None
        eng.forward(loc[0], loc[1], loc[2], loc[3], loc[4], max_qlen, max_kv_len, loc[5], local, want_logits, 0)
        if n_rows:
            if logits_out is None:
                with torch.cuda.device(src):
                    logits_out = torch.empty(n_rows, V, dtype=torch.float32, device=src)
            dst = logits_out if logits_out.dim() == 2 else logits_out.view(-1, V)
            dst.copy_(local, non_blocking=True)          # peer copy over NVLink, strided destinations welcome
            logits_out.record_stream(side)
        done = torch.cuda.Event()
        done.record(side)
    torch.cuda.current_stream(src).wait_event(done)
    return logits_out if n_rows else None


class TPQwenEngine(_BatchOps):
    """Tensor-parallel target inside ONE process: rank r of the Megatron split is a ``QwenEngine`` on
    ``gpu_ids[r]``; the row-parallel boundaries exchange partial sums through peer-mapped memory
    (``asd_engine_peer_connect``: cudaDeviceEnablePeerAccess, no IPC, no torchrun).  This is what
    ``Stage(tensor_parallel_size=t, gpu_ids=[...])`` builds, mirroring how the reference hands
    ``tensor_parallel_size`` to vLLM (src/serving/real_model_pipeline.py:98-108, configs/qwen3_models.yaml:10-51).
    Same calling interface as ``QwenEngine``; every rank's launches are issued by its own host thread on its
    own stream, logits come from rank 0 (lm_head is replicated, all ranks compute identical logits)."""

    def __init__(self, cfg: Qwen2Config, gpu_ids, max_seqs: int, max_seq_len: int, max_tokens: int = 256,
                 page_size: int = 16, shuffle_pages: bool = True, fuse_norm: bool = True):
        from concurrent.futures import ThreadPoolExecutor
        gpu_ids = [int(g) for g in gpu_ids]
        if len(set(gpu_ids)) != len(gpu_ids) or len(gpu_ids) < 2:
            raise AsdError(f"tensor parallelism needs >= 2 distinct GPUs, got {gpu_ids}")
        if max(gpu_ids) >= torch.cuda.device_count():
            raise AsdError(f"gpu_ids {gpu_ids} but only {torch.cuda.device_count()} CUDA devices are visible")
        t = len(gpu_ids)
        self.cfg, self.gpu_ids, self.tp_size = cfg, gpu_ids, t
        self.ranks = [QwenEngine(cfg, max_seqs, max_seq_len, max_tokens, page_size, tp_rank=r, tp_size=t,
                                 device=f"cuda:{g}", shuffle_pages=shuffle_pages, fuse_norm=fuse_norm)
                      for r, g in enumerate(gpu_ids)]
        self.device = self.ranks[0].device
        self.max_seqs, self.max_seq_len, self.max_tokens, self.page_size = max_seqs, max_seq_len, max_tokens, page_size
        self.max_pages = self.ranks[0].max_pages
        arr = (ctypes.c_void_p * t)(*[e.h for e in self.ranks])
        check(lib().asd_engine_peer_connect(arr, t), "asd_engine_peer_connect")
        self._pool = ThreadPoolExecutor(max_workers=t, thread_name_prefix="asd-tp")
        self._lock = threading.Lock()
        self._index_cache = {}

    # QwenEngine's uniform-batch helpers work on anything with forward()/device/_index_cache
    def close(self):
        for e in self.ranks:
            e.close()
        self._pool.shutdown(wait=False)

    def load_hf_weights(self, w):
        for e in self.ranks:
            e.load_hf_weights(w)
        return self

    def load_random(self, seed: int, std: float = 0.02):
        for e in self.ranks:
            e.load_random(seed, std)
        return self

    def set_option(self, name: str, value: int):
        for e in self.ranks:
            e.set_option(name, value)

    def tp_error(self) -> int:
        return max(e.tp_error() for e in self.ranks)

    def profile_read(self):
        return self.ranks[0].profile_read()

    def _rank_forward(self, r, ready, src, ins, max_qlen, max_kv_len, n_rows, want_logits, logits_out):
        eng = self.ranks[r]
        side = eng._hop_stream()
        with torch.cuda.device(eng.device), torch.cuda.stream(side):
            side.wait_event(ready)
            loc = []
            for t in ins:
                if t is None:
                    loc.append(None)
                elif t.device == eng.device:
                    loc.append(t)
                    t.record_stream(side)
                else:
                    loc.append(t.to(eng.device, non_blocking=True))
                    t.record_stream(side)
            want = want_logits and r == 0          # replicated lm_head: only rank 0's copy is consumed
            local = None
            if want and n_rows:
                direct = logits_out is not None and logits_out.device == eng.device and logits_out.dim() == 2 \
                    and logits_out.stride(1) == 1
                local = logits_out if direct else eng._hop_logits(n_rows)
            ld = local.stride(0) if (local is not None and local is logits_out) else 0
            eng.forward(loc[0], loc[1], loc[2], loc[3], loc[4], max_qlen, max_kv_len, loc[5], local, want, ld)
            if want and n_rows and local is not logits_out:
                dst = logits_out if logits_out.dim() == 2 else logits_out.view(-1, self.cfg.vocab_size)
                dst.copy_(local, non_blocking=True)
            if want and n_rows:
                logits_out.record_stream(side)
            done = torch.cuda.Event()
            done.record(side)
        return done

    def forward(self, tokens, positions, token_slot, cu_q, seq_slot, max_qlen: int, max_kv_len: int,
                logit_rows: Optional[torch.Tensor] = None, logits_out: Optional[torch.Tensor] = None,
                want_logits: bool = True, logits_ld: int = 0):
        src = tokens.device
        n_rows = 0 if not want_logits else (tokens.numel() if logit_rows is None else logit_rows.numel())
        if n_rows and logits_out is None:
            with torch.cuda.device(src):
                logits_out = torch.empty(n_rows, self.cfg.vocab_size, dtype=torch.float32, device=src)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(src))
        ins = (tokens, positions, token_slot, cu_q, seq_slot, logit_rows)
        with self._lock:
            futs = [self._pool.submit(self._rank_forward, r, ready, src, ins, max_qlen, max_kv_len, n_rows,
                                      want_logits, logits_out) for r in range(self.tp_size)]
            dones = [f.result() for f in futs]
        cur = torch.cuda.current_stream(src)
        for d in dones:     # the next step's inputs are produced on `cur`: keep every rank behind it in lock step
            cur.wait_event(d)
        return logits_out if n_rows else None


class SpecDecoder:
    """Chain draft-then-verify: k draft steps on the draft engine, ONE (k+1)-token verify forward on
    the target engine, ONE fused rejection-sampling launch; state stays on the device.

    ``draft=None`` degenerates to plain autoregressive decoding of the target (k = 0).  The draft engine may
    live on another GPU than the target (cascade stages on disjoint GPU sets): its forwards hop over NVLink.
    ``set_limit(n)`` freezes a sequence once it has emitted n tokens (its state stops advancing, so positions
    stay below prompt + n + k + 1 however long the slowest sequence of the batch takes)."""

    def __init__(self, target, draft, batch: int, k: int, temperature: float, seed: int = 4321):
        self.t, self.d = target, draft
        self.B, self.k, self.T = batch, (k if draft is not None else 0), float(temperature)
        dev = target.device
        self.device = dev
        V = target.cfg.vocab_size
        if draft is not None and draft.cfg.vocab_size != V:
            raise AsdError("draft and target must share the vocabulary")
        if batch > target.max_seqs or (draft is not None and batch > draft.max_seqs):
            raise AsdError(f"batch {batch} exceeds the engines' sequence slots")
        from .ops import RejectionSampler
        with torch.cuda.device(dev):
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
            self.emitted = torch.zeros(batch, dtype=torch.int32, device=dev)
            self._empty_i = torch.zeros(batch, 0, dtype=torch.int32, device=dev)
            self._empty_d = torch.zeros(batch, 0, dtype=torch.float64, device=dev)
        self.kv_bound = 0     # host-side upper bound of every sequence length (no sync needed)
        self.limit = None     # tokens per sequence after which it is frozen (set_limit)
        self.start_len = 0
        self.max_len = min(target.max_seq_len, draft.max_seq_len) if draft is not None else target.max_seq_len
        self.time_verify = False
        self.verify_events = []

    def _uniform(self, *shape):
        return torch.rand(*shape, dtype=torch.float64, device=self.device, generator=self.gen)

    def _sample_rows(self, logits_bv: torch.Tensor) -> torch.Tensor:
        """one token per row from softmax(logits/T) (argmax when T <= 0), fused kernel with k = 0"""
        out = self.row_sampler(logits_bv, None, self._empty_i, self._empty_d, self._uniform(self.B), self.T)
        return out["out_tokens"][:, 0].clone()

    def set_limit(self, max_new_tokens: int):
        self.limit = int(max_new_tokens)

    def _begin(self, lens: torch.Tensor, max_len: int):
        self.pos = lens.to(torch.int32).to(self.device).contiguous()
        self.emitted = torch.ones(self.B, dtype=torch.int32, device=self.device)    # the token sampled by prefill
        self.start_len = max_len
        self.kv_bound = max_len + 1

    def prefill(self, prompt_ids):
        """prompt_ids: int tensor [B, P] (equal lengths) or a list of B token-id lists (ragged).  Fills both KV
        caches and samples the first token."""
        with torch.cuda.device(self.device):
            if isinstance(prompt_ids, (list, tuple)):
                assert len(prompt_ids) == self.B
                lens = [len(p) for p in prompt_ids]
                if max(lens) + 1 > self.max_len:
                    raise AsdError(f"prompt of {max(lens)} tokens does not fit max_seq_len {self.max_len}")
                logits = self.t.prefill_ragged(prompt_ids, self.slots, device=self.device)
                if self.d is not None:
                    self.d.prefill_ragged(prompt_ids, self.slots, device=self.device)
                self.last_tok = self._sample_rows(logits.contiguous())
                self.prev_tok = torch.tensor([p[-1] for p in prompt_ids], dtype=torch.int32).to(self.device)
                self._begin(torch.tensor(lens, dtype=torch.int32), max(lens))
                return self.last_tok
            B, P = prompt_ids.shape
            assert B == self.B and P >= 1
            if P + 1 > self.max_len:
                raise AsdError(f"prompt of {P} tokens does not fit max_seq_len {self.max_len}")
            prompt_ids = prompt_ids.to(self.device)
            logits = self.t.prefill(prompt_ids, self.slots)
            if self.d is not None:
                self.d.prefill(prompt_ids, self.slots, want_logits=False)
            self.last_tok = self._sample_rows(logits.contiguous())
            self.prev_tok = prompt_ids[:, -1].to(torch.int32).contiguous()
            self._begin(torch.full((B,), P, dtype=torch.int32), P)
            return self.last_tok

    def seed_state(self, prefix_len: int, last_tok: torch.Tensor, prev_tok: torch.Tensor):
        """Benchmark helper: declare that both KV caches already hold ``prefix_len`` positions."""
        self.last_tok, self.prev_tok = last_tok.to(torch.int32).contiguous(), prev_tok.to(torch.int32).contiguous()
        self._begin(torch.full((self.B,), prefix_len, dtype=torch.int32), prefix_len)

    def step(self):
        """One draft-then-verify step for the whole batch.  Returns the sampler's output dict
        (device tensors): out_tokens [B, k+1] (-1 padded), accepted_len [B], accept_mask, features."""
        with torch.cuda.device(self.device):
            return self._step()

    def _step(self):
        B, k, V = self.B, self.k, self.t.cfg.vocab_size
        bound = self.kv_bound + k + 1
        if self.limit is not None:      # frozen sequences stop advancing: positions stay below this
            bound = min(bound, self.start_len + self.limit + k + 1)
        if bound > self.max_len:
            # the host-side bound assumes every draft token of every step was accepted; before refusing, read the
            # true longest sequence (one sync, only on this rare path) and tighten the bound
            bound = int(self.pos.max().item()) + k + 1
            if bound > self.max_len:
                raise AsdError(f"step would reach position {bound} but max_seq_len is {self.max_len} "
                               "(raise max_model_len or lower max_tokens)")
        if k > 0:
            # draft step 1 re-feeds the previous token so the draft KV is complete after an all-accept
            two = torch.stack([self.prev_tok, self.last_tok], 1)
            self.d.forward_uniform(two, self.pos - 1, self.slots, bound, last_only=True,
                                   logits_out=self.draft_logits[:, 0], logits_ld=k * V)
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
        self.t.forward_uniform(toks, self.pos, self.slots, bound, logits_out=self.target_logits.view(B * (k + 1), V))
        if self.time_verify:
            ev[1].record()
            self.verify_events.append(ev)
        out = self.sampler(self.target_logits, self.draft_logits if k > 0 and self.T > 0 else None,
                           self.draft_tokens[:, :k].contiguous() if k > 0 else self._empty_i,
                           self._uniform(B, k) if k > 0 else self._empty_d, self._uniform(B), self.T)
        n = out["accepted_len"].to(torch.int64)
        new_last = out["out_tokens"].gather(1, n[:, None])[:, 0]
        new_prev = toks.gather(1, n[:, None])[:, 0]
        adv = out["accepted_len"] + 1
        if self.limit is not None:
            live = self.emitted < self.limit
            out["live"] = live
            new_last = torch.where(live, new_last, self.last_tok)
            new_prev = torch.where(live, new_prev, self.prev_tok)
            adv = adv * live.to(torch.int32)
        self.prev_tok = new_prev.contiguous()
        self.last_tok = new_last.contiguous()
        self.pos = self.pos + adv
        self.emitted = self.emitted + adv
        self.kv_bound = bound
        return out

    def step_host(self, host_state: torch.Tensor, host_tokens: torch.Tensor, host_accepted: torch.Tensor):
        """Same step driven from HOST (pinned) buffers, the way a non-torch integrator would call it:
        host_state int32 [3, B] = (last_tok, prev_tok, pos) is copied to the device, the step runs, then
        the emitted tokens [B, k+1], accepted lengths [B] and the next state (in place) are copied back;
        returns after the copy-out has completed."""
        with torch.cuda.device(self.device):
            dev_state = host_state.to(self.device, non_blocking=True)
            self.last_tok, self.prev_tok, self.pos = (dev_state[0].contiguous(), dev_state[1].contiguous(),
                                                      dev_state[2].contiguous())
            out = self._step()
            host_tokens.copy_(out["out_tokens"], non_blocking=True)
            host_accepted.copy_(out["accepted_len"], non_blocking=True)
            host_state.copy_(torch.stack([self.last_tok, self.prev_tok, self.pos]), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for e in (self.t, self.d):
            if e is not None and getattr(e, "tp_size", 1) > 1 and e.tp_error():
                raise AsdError("tensor-parallel peer did not answer within the spin bound (asd_engine_tp_error)")
        return host_tokens, host_accepted

    def generate(self, prompts, max_new_tokens: int):
        """Prefill + decode until every sequence has ``max_new_tokens`` tokens.  Everything a step emits is
        appended to device buffers; the host only reads the shortest sequence's length, one step late (the
        read of step s overlaps step s + 1), so the GPU never waits for the host.
        Returns (tokens [B, n] int32, logprobs [B, n] fp32, features [B, n, F] fp32, stats) on the host."""
        B, k = self.B, self.k
        with torch.cuda.device(self.device):
            self.set_limit(max_new_tokens)
            first = self.prefill(prompts)
            cap = max_new_tokens + k + 2
            toks = torch.full((B, cap), -1, dtype=torch.int32, device=self.device)
            lps = torch.zeros(B, cap, dtype=torch.float32, device=self.device)
            feats = torch.zeros(B, cap, NUM_FEATURES, dtype=torch.float32, device=self.device)
            toks[:, 0] = first
            col = torch.arange(k + 1, device=self.device)[None]
            accepted = torch.zeros((), dtype=torch.int64, device=self.device)
            pend = []           # (event, pinned flag) of earlier steps
            steps = 0
            done = max_new_tokens <= 1
            while not done:
                at = self.emitted.to(torch.int64)[:, None] + col        # where this step's tokens go
                out = self._step()
                steps += 1
                idx = at.clamp(max=cap - 1)
                keep = (col <= out["accepted_len"].to(torch.int64)[:, None]) & out["live"][:, None]
                toks.scatter_(1, idx, torch.where(keep, out["out_tokens"], toks.gather(1, idx)))
                lps.scatter_(1, idx, torch.where(keep, out["out_logprobs"], lps.gather(1, idx)))
                fi = idx[:, :, None].expand(-1, -1, NUM_FEATURES)
                feats.scatter_(1, fi, torch.where(keep[:, :, None], out["features"], feats.gather(1, fi)))
                accepted += (out["accepted_len"].to(torch.int64) * out["live"]).sum()
                flag = torch.empty(1, dtype=torch.int32).pin_memory()
                flag.copy_(self.emitted.min().reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pend.append((ev, flag))
                if len(pend) > 1:           # look at the step BEFORE the one just enqueued
                    ev0, f0 = pend.pop(0)
                    ev0.synchronize()
                    done = int(f0.item()) >= max_new_tokens
            torch.cuda.current_stream().synchronize()
            n = max_new_tokens
            feats_dev = feats[:, :n].contiguous()      # stays on the device for asd_cascade_decide (scorer -> stop rule)
            res = (toks[:, :n].cpu(), lps[:, :n].cpu(), feats_dev.cpu(),
                   {"accepted": int(accepted.item()), "steps": steps, "features_device": feats_dev})
        for e in (self.t, self.d):
            if e is not None and getattr(e, "tp_size", 1) > 1 and e.tp_error():
                raise AsdError("tensor-parallel peer did not answer within the spin bound (asd_engine_tp_error)")
        return res

    def _sample_rows_strided(self, i: int) -> torch.Tensor:
        # the fused sampler wants contiguous rows: draft step i wrote rows with stride k*V
        return self._sample_rows(self.draft_logits[:, i].contiguous())
