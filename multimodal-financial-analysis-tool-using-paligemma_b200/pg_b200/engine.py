"""Host side of the B200 PaliGemma engine: weight repack, paged KV pool, and the launch
sequences (vision tower, prefill / cache-off recompute, decode step, CUDA-graph decode loop).

PyTorch is used for device memory, streams and CUDA-graph capture only; every FLOP of the
hot path runs in libpg_b200.so through the C ABI of include/pg_b200.h.  There is no CPU or
PyTorch-op fallback: a missing library or a non-CUDA tensor raises.

Reference semantics reproduced: attention is never masked (Q1);
prefill positions 0..N-1 (Q2); cached steps use position = attention-mask length, i.e.
N+t, skipping N (Q3); a q_len>1 forward on a non-empty cache gives every new token that
same position (Q6, patched merge); pad ids embed to zero rows (Q8); the sqrt(D)
normaliser is rounded to the model dtype while image features are divided by the exact
float (Q9); RoPE inv_freq is rounded to the model dtype by `model.to(dtype)`.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional

import torch

from . import _cabi as cabi
from ._cabi import exref, ptr
from .dist import TP, check_divisible, combine_argmax_keys, shard_batch, shard_rows, shard_text_layer


_WORKSPACES: Dict[tuple, torch.Tensor] = {}   # per device, kept for the life of the process (pg_set_workspace)


def _get(obj, name, default=None):
    if isinstance(obj, Mapping):
        return obj.get(name, default)
    return getattr(obj, name, default)


@dataclass
class Dims:
    # text
    V: int; D: int; F: int; L: int; nq: int; nkv: int; hd: int; eps: float; theta: float; max_pos: int
    # vision
    Hv: int; Iv: int; Lv: int; heads_v: int; C: int; S: int; p: int; eps_v: float
    # glue
    image_token_index: int; pad_token_id: int; hidden_size: int

    @property
    def P(self) -> int:
        return (self.S // self.p) ** 2

    @staticmethod
    def from_config(cfg) -> "Dims":
        v, t = _get(cfg, "vision_config"), _get(cfg, "text_config")
        pad = _get(cfg, "pad_token_id")
        return Dims(
            V=_get(t, "vocab_size"), D=_get(t, "hidden_size"), F=_get(t, "intermediate_size"),
            L=_get(t, "num_hidden_layers"), nq=_get(t, "num_attention_heads"),
            nkv=_get(t, "num_key_value_heads"), hd=_get(t, "head_dim", 256),
            eps=_get(t, "rms_norm_eps", 1e-6), theta=_get(t, "rope_theta", 10000.0),
            max_pos=_get(t, "max_position_embeddings", 8192),
            Hv=_get(v, "hidden_size"), Iv=_get(v, "intermediate_size"), Lv=_get(v, "num_hidden_layers"),
            heads_v=_get(v, "num_attention_heads"), C=_get(v, "num_channels", 3), S=_get(v, "image_size"),
            p=_get(v, "patch_size"), eps_v=_get(v, "layer_norm_eps", 1e-6),
            image_token_index=_get(cfg, "image_token_index"), pad_token_id=-1 if pad is None else pad,
            hidden_size=_get(cfg, "hidden_size"))


class PagedKV:
    """Device state of one batch of sequences in the paged KV pool (the storage behind the
    reference's KVCache, modeling_gemma.py:10-36)."""

    MIN_TABLE_PAGES = 256                    # table columns allocated up front (16 K tokens at 64 per page)

    def __init__(self, engine: "PaliGemmaEngine", batch: int):
        self.engine = engine
        self.batch = batch
        self.length = 0                      # entries per sequence (host mirror)
        self.pages: List[List[int]] = [[] for _ in range(batch)]
        self.max_pages = 0
        self.page_table: Optional[torch.Tensor] = None   # int32 [batch, max_pages] device
        self.kv_len = torch.zeros(batch, dtype=torch.int32, device=engine.device)

    def reserve(self, total_len: int) -> None:
        """Make sure every sequence owns pages for `total_len` entries.  The device page table is allocated once with
        room for MIN_TABLE_PAGES pages per sequence and filled in place as pages are added, so its address and row
        stride never change while a sequence grows: a decode graph captured over it stays valid."""
        eng = self.engine
        need = (total_len + eng.page_size - 1) // eng.page_size
        have = len(self.pages[0])
        if need <= have:
            return
        grow = max(need, min(2 * have, need + 8))
        if self.page_table is None or grow > self.max_pages:
            cap = max(self.MIN_TABLE_PAGES, 2 * grow)
            table = torch.zeros((self.batch, cap), dtype=torch.int32, device=eng.device)
            if self.page_table is not None and have:
                table[:, :have].copy_(self.page_table[:, :have])
            self.page_table, self.max_pages = table, cap
        new = [eng._alloc_pages(grow - have) for _ in range(self.batch)]
        for b in range(self.batch):
            self.pages[b].extend(new[b])
        self.page_table[:, have:grow].copy_(torch.tensor(new, dtype=torch.int32))

    def release(self) -> None:
        if self.engine is not None:
            for p in self.pages:
                self.engine._free_pages(p)
        self.pages = [[] for _ in range(self.batch)]
        self.length = 0
        self.page_table = None
        self.max_pages = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def gather(self, layer: int, which: str) -> torch.Tensor:
        """Contiguous (B, n_kv, T, hd) copy of one layer's keys or values."""
        eng, d = self.engine, self.engine.dims
        out = torch.empty((self.batch, d.nkv, self.length, d.hd), dtype=eng.dtype, device=eng.device)
        if self.length:
            pool = eng.k_pool[layer] if which == "k" else eng.v_pool[layer]
            cabi.check(cabi.lib().pg_kv_gather(ptr(out), ptr(pool), ptr(self.page_table), self.max_pages,
                                               eng.page_size, self.batch, self.length, d.nkv, d.hd,
                                               eng.dt, cabi.stream()), "kv_gather")
        return out


class PaliGemmaEngine:
    """Owns repacked weights + the KV pool and issues the kernel sequences."""

    def __init__(self, config, weights: Mapping[str, torch.Tensor], *, device=None, dtype=None,
                 page_size: int = 64, kv_pool_tokens: int = 65536, gemm_impl: int = 0,
                 adopt=None, tp=None, batched_min: Optional[int] = None):
        cabi.lib()  # fail now if the CUDA library is missing
        self.dims = d = Dims.from_config(config)
        self.tp = tp if tp is not None else TP()
        if self.tp.active:
            check_divisible(d, self.tp.size)
            adopt = None  # shards do not have the parameters' shapes
        self.nq_l, self.F_l, self.V_l = d.nq // self.tp.size, d.F // self.tp.size, d.V // self.tp.size
        emb = weights["language_model.model.embed_tokens.weight"]
        self.device = torch.device(device) if device is not None else emb.device
        if self.device.type != "cuda":
            raise RuntimeError("pg_b200 runs on CUDA devices only (no CPU fallback); move the model to 'cuda'")
        self.dtype = dtype or emb.dtype
        if self.dtype not in cabi.DTYPE_CODE:
            raise RuntimeError(f"unsupported model dtype {self.dtype}")
        self.dt = cabi.DTYPE_CODE[self.dtype]
        self.gemm_impl = gemm_impl
        # rows from which the decode step / last-position lm_head use the tensor-core GEMMs instead of the GEMV kernels
        # (fp32 verification mode has no tensor-core path: GEMV up to its 8-row limit)
        self.batched_min = cabi.BATCHED_DECODE_MIN if self.dtype != torch.float32 else cabi.MAX_DECODE_BATCH + 1
        if batched_min is not None:     # tests: drive the batched step's launch sequence in fp32 (SIMT GEMMs) too
            self.batched_min = int(batched_min)
        self._op_timing = os.environ.get("PG_OP_TIMING", "0") == "1"
        self._vision_graphs_on = os.environ.get("PG_VISION_GRAPH", "1") != "0"
        self._vision_graphs: Dict[tuple, tuple] = {}
        self._op_events = []
        self.page_size = page_size
        self._vec = 4 if self.dtype == torch.float32 else 8
        # tcgen05 attention (16-bit dtypes): SigLIP heads padded with zero columns; Gemma needs hd 256 and 64-token pages
        tc_ok = self.dtype != torch.float32 and os.environ.get("PG_ATTN_TC", "1") != "0"
        hdv = d.Hv // d.heads_v
        # the ViT attention kernel multiplies 16-column steps (72 -> 80); the tiled kernel needs 64-column blocks (-> 128)
        vit = hdv <= 80 and d.P <= 256 and os.environ.get("PG_ATTN_VIT", "1") != "0"
        self.vision_head_pad = ((hdv + 15) // 16) * 16 if vit else 128
        self.attn_tc_vision = tc_ok and hdv <= 128 and hdv % 8 == 0
        self.attn_tc_text = tc_ok and d.hd == 256 and page_size == 64
        if self.attn_tc_vision:
            adopt_outer = adopt
            adopt = (lambda k, v: None if ".vision_model." in k else adopt_outer(k, v)) if adopt_outer else None
        self._repack(weights, adopt)
        # paged KV pool: [L, pages, page_size, nkv*hd] for K and for V
        self.num_pages = max(8, (kv_pool_tokens + page_size - 1) // page_size)
        self.k_pool = torch.zeros((d.L, self.num_pages, page_size, d.nkv * d.hd), dtype=self.dtype, device=self.device)
        self.v_pool = torch.zeros_like(self.k_pool)
        self._free = list(range(self.num_pages - 1, -1, -1))
        # pinned, device-mapped: kernels store to it only when they meet a bad id, the host reads it without a sync
        self.err_flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._err_np = self.err_flag.numpy()
        self.max_splits = 32
        self._decode_states: Dict[tuple, "DecodeState"] = {}
        # scratch for the split-K prompt GEMMs (fp32 partials; 512 tokens x 2F features is the largest user)
        if self.dtype != torch.float32:
            key = (self.device.index, )
            if key not in _WORKSPACES:
                _WORKSPACES[key] = torch.empty(int(os.environ.get("PG_WORKSPACE_MB", "80")) << 20, dtype=torch.uint8,
                                               device=self.device)
            ws = _WORKSPACES[key]
            cabi.check(cabi.lib().pg_set_workspace(ptr(ws), ws.numel()), "set_workspace")
        # tensor parallel: the peer-memory exchange the decode kernels use instead of collectives (dist.Fabric)
        self.fabric = self.tp.make_fabric(d.D, self.device)
        if self.fabric is not None and 2 * d.L + 2 > 4096:
            raise ValueError("too many layers for the exchange sequence numbering")

    def kv_pool_bytes(self) -> int:
        return 2 * self.k_pool.numel() * self.k_pool.element_size()

    def kv_bytes_in_use(self) -> int:
        """Bytes of the pre-allocated pool that hold pages of live sequences (the reference's KVCache grows by
        torch.cat, so its allocator peak tracks this number; ours pre-allocates the pool)."""
        d = self.dims
        used = self.num_pages - len(self._free)
        return 2 * d.L * used * self.page_size * d.nkv * d.hd * self.k_pool.element_size()

    # ------------------------------------------------------------------ weights
    def _w(self, weights, key):
        t = weights[key]
        if t.device != self.device or t.dtype != self.dtype:
            t = t.to(device=self.device, dtype=self.dtype)
        return t.contiguous()

    def _repack(self, weights, adopt):
        """Fuse q/k/v and gate/up, pad the patch-embed K, keep HF layout otherwise.
        `adopt(key, view)` lets the nn.Module owner re-point its parameters at the fused
        storage so the checkpoint is not held twice."""
        d = self.dims
        W = lambda k: self._w(weights, k)
        vm = "vision_tower.vision_model."
        kc = d.C * d.p * d.p
        self.k_patch = ((kc + self._vec - 1) // self._vec) * self._vec
        wp = torch.zeros((d.Hv, self.k_patch), dtype=self.dtype, device=self.device)
        wp[:, :kc] = W(vm + "embeddings.patch_embedding.weight").reshape(d.Hv, kc)
        self.v_patch_w, self.v_patch_b = wp, W(vm + "embeddings.patch_embedding.bias")
        self.v_pos = W(vm + "embeddings.position_embedding.weight")
        self.v_layers = []
        for i in range(d.Lv):
            Lk = f"{vm}encoder.layers.{i}."
            qkv_w = torch.cat([W(Lk + f"self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj")], 0)
            qkv_b = torch.cat([W(Lk + f"self_attn.{n}.bias") for n in ("q_proj", "k_proj", "v_proj")], 0)
            if adopt:
                for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
                    adopt(Lk + f"self_attn.{n}.weight", qkv_w[j * d.Hv:(j + 1) * d.Hv])
            if self.attn_tc_vision:
                # head rows padded 72 -> 80 (or 128) with zero weights / zero bias: the tcgen05 attention kernels
                # multiply whole 16-column steps (64-column blocks), and zeros contribute nothing to QK^T or PV
                hdv, hp, nh = d.Hv // d.heads_v, self.vision_head_pad, d.heads_v
                wp_ = torch.zeros((3, nh, hp, d.Hv), dtype=self.dtype, device=self.device)
                wp_[:, :, :hdv] = qkv_w.view(3, nh, hdv, d.Hv)
                bp_ = torch.zeros((3, nh, hp), dtype=self.dtype, device=self.device)
                bp_[:, :, :hdv] = qkv_b.view(3, nh, hdv)
                qkv_w, qkv_b = wp_.view(3 * nh * hp, d.Hv), bp_.view(-1)
            self.v_layers.append(dict(
                ln1_w=W(Lk + "layer_norm1.weight"), ln1_b=W(Lk + "layer_norm1.bias"),
                qkv_w=qkv_w, qkv_b=qkv_b,
                o_w=W(Lk + "self_attn.out_proj.weight"), o_b=W(Lk + "self_attn.out_proj.bias"),
                ln2_w=W(Lk + "layer_norm2.weight"), ln2_b=W(Lk + "layer_norm2.bias"),
                fc1_w=W(Lk + "mlp.fc1.weight"), fc1_b=W(Lk + "mlp.fc1.bias"),
                fc2_w=W(Lk + "mlp.fc2.weight"), fc2_b=W(Lk + "mlp.fc2.bias")))
        self.v_post_w, self.v_post_b = W(vm + "post_layernorm.weight"), W(vm + "post_layernorm.bias")
        self.proj_w, self.proj_b = W("multi_modal_projector.linear.weight"), W("multi_modal_projector.linear.bias")
        lm = "language_model.model."
        self.emb = W(lm + "embed_tokens.weight")
        head = weights.get("language_model.lm_head.weight")
        self.lm_head = self.emb if (head is None or head.data_ptr() == weights[lm + "embed_tokens.weight"].data_ptr()) \
            else W("language_model.lm_head.weight")
        if self.tp.active:
            self.lm_head = shard_rows(self.lm_head, self.tp.rank, self.tp.size)
        self.t_layers = []
        for i in range(d.L):
            Lk = f"{lm}layers.{i}."
            names = ("q_proj", "k_proj", "v_proj")
            if self.tp.active:
                qkv, o, gu, down = shard_text_layer(
                    W(Lk + "self_attn.q_proj.weight"), W(Lk + "self_attn.k_proj.weight"), W(Lk + "self_attn.v_proj.weight"),
                    W(Lk + "self_attn.o_proj.weight"), W(Lk + "mlp.gate_proj.weight"), W(Lk + "mlp.up_proj.weight"),
                    W(Lk + "mlp.down_proj.weight"), self.tp.rank, self.tp.size)
            else:
                qkv = torch.cat([W(Lk + f"self_attn.{n}.weight") for n in names], 0)
                gu = torch.cat([W(Lk + "mlp.gate_proj.weight"), W(Lk + "mlp.up_proj.weight")], 0)
                o, down = W(Lk + "self_attn.o_proj.weight"), W(Lk + "mlp.down_proj.weight")
            if adopt:
                off = 0
                for n in names:
                    rows = weights[Lk + f"self_attn.{n}.weight"].shape[0]
                    adopt(Lk + f"self_attn.{n}.weight", qkv[off:off + rows])
                    off += rows
                adopt(Lk + "mlp.gate_proj.weight", gu[:d.F])
                adopt(Lk + "mlp.up_proj.weight", gu[d.F:])
            self.t_layers.append(dict(
                ln1=W(Lk + "input_layernorm.weight"), qkv=qkv, o=o,
                ln2=W(Lk + "post_attention_layernorm.weight"), gu=gu, down=down))
        self.final_norm = W(lm + "norm.weight")
        # RoPE frequencies (modeling_gemma.py:151): the buffer is rounded by model.to(dtype)
        f = 1.0 / (d.theta ** (torch.arange(0, d.hd, 2, dtype=torch.int64).float() / d.hd))
        self.inv_freq = f.to(self.dtype).float().to(self.device)
        # sqrt(D) normaliser rounded to the model dtype (:367); image divisor is the exact float (:481)
        self.normalizer = float(torch.tensor(d.D ** 0.5, dtype=self.dtype).float())
        self.img_div = float(torch.tensor(d.hidden_size ** 0.5, dtype=torch.float32))

    def weight_bytes_per_decode_step(self) -> int:
        """Algorithmic HBM bytes one decode step must read from the weights."""
        d, e = self.dims, torch.tensor([], dtype=self.dtype).element_size()
        per_layer = ((self.nq_l + 2 * d.nkv) * d.hd * d.D + d.D * self.nq_l * d.hd + 3 * self.F_l * d.D + 2 * d.D)
        return e * (d.L * per_layer + self.V_l * d.D + d.D)

    # ------------------------------------------------------------------ KV pages
    def _alloc_pages(self, n: int) -> List[int]:
        if n > len(self._free):
            raise RuntimeError(f"KV pool exhausted: need {n} pages, {len(self._free)} free "
                               f"(pool holds {self.num_pages * self.page_size} tokens; raise kv_pool_tokens)")
        out = [self._free.pop() for _ in range(n)]
        return out

    def _free_pages(self, pages: List[int]) -> None:
        self._free.extend(reversed(pages))

    def new_kv(self, batch: int) -> PagedKV:
        return PagedKV(self, batch)

    # ------------------------------------------------------------------ op wrappers
    def _gemm(self, out, a, w, bias=None, res=None, epi=cabi.EPI_NONE, res_mod=0, out_f32=False,
              M=None, N=None, K=None, lda=None, ldc=None):
        M = a.shape[0] if M is None else M
        K = a.shape[1] if K is None else K
        if N is None:
            N = w.shape[0] // 2 if epi == cabi.EPI_GEGLU else w.shape[0]
        lda = a.stride(0) if lda is None else lda
        ldc = out.stride(0) if ldc is None else ldc
        cabi.check(cabi.lib().pg_gemm(ptr(out), ptr(a), ptr(w), ptr(bias), ptr(res), M, N, K, lda,
                                      w.stride(0), ldc, 0 if res is None else res.stride(0), res_mod, epi,
                                      1 if out_f32 else 0, self.gemm_impl, self.dt, cabi.stream()), "gemm")
        return out

    def _new(self, *shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.dtype, device=self.device)

    # ------------------------------------------------------------------ per-op timing (diagnostic: PG_OP_TIMING=1)
    class _NoTimer:
        def __enter__(self): return self
        def __exit__(self, *a): return False

    class _EvTimer:
        def __init__(self, sink, name): self.sink, self.name = sink, name
        def __enter__(self):
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
            return self
        def __exit__(self, *a):
            self.e1.record()
            self.sink.append((self.name, self.e0, self.e1))
            return False

    def _op_timer(self, name):
        """CUDA-event bracket around one op when PG_OP_TIMING=1 (in-pipeline durations, L2 state as in the real
        run); `op_times()` sums them per op name.  A no-op otherwise."""
        if not self._op_timing:
            return self._NoTimer()
        return self._EvTimer(self._op_events, name)

    def op_times(self, reset: bool = True) -> Dict[str, float]:
        torch.cuda.synchronize()
        out: Dict[str, float] = {}
        for name, e0, e1 in self._op_events:
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        if reset:
            self._op_events = []
        return out

    # ------------------------------------------------------------------ vision tower + projector
    def vision_features(self, pixels: torch.Tensor) -> torch.Tensor:
        """SiglipVisionModel.forward (modeling_siglip.py:236-255): (B,C,S,S) -> (B,P,Hv)."""
        d, L = self.dims, cabi.lib()
        if pixels.device != self.device:
            raise RuntimeError("pixel_values must live on the model's CUDA device")
        px = pixels.to(self.dtype).contiguous()
        B, T = px.shape[0], px.shape[0] * d.P
        st = cabi.stream()
        col = self._new(T, self.k_patch)
        cabi.check(L.pg_im2col(ptr(col), ptr(px), B, d.C, d.S, d.S, d.p, self.k_patch, self.dt, st), "im2col")
        h = self._new(T, d.Hv)
        self._gemm(h, col, self.v_patch_w, self.v_patch_b, self.v_pos, cabi.EPI_BIAS_RES, res_mod=d.P)
        hdv = d.Hv // d.heads_v
        qkv_cols = 3 * d.heads_v * self.vision_head_pad if self.attn_tc_vision else 3 * d.Hv
        ln, qkv, att = self._new(T, d.Hv), self._new(T, qkv_cols), self._new(T, d.Hv)
        h2, mid = self._new(T, d.Hv), self._new(T, d.Iv)
        scale = float(hdv ** -0.5)
        tm = self._op_timer
        for w in self.v_layers:
            with tm("ln"):
                cabi.check(L.pg_layernorm(ptr(ln), ptr(h), ptr(w["ln1_w"]), ptr(w["ln1_b"]), T, d.Hv, d.eps_v, self.dt, st), "ln1")
            with tm("qkv"):
                self._gemm(qkv, ln, w["qkv_w"], w["qkv_b"], None, cabi.EPI_BIAS)
            with tm("attention"):
                if self.attn_tc_vision:
                    hp = self.vision_head_pad
                    cabi.check(L.pg_attention_tc(ptr(att), d.Hv, ptr(qkv), T, qkv_cols, 0, ptr(qkv), ptr(qkv), T, qkv_cols,
                                                 d.heads_v * hp, 2 * d.heads_v * hp, hp, d.P, None, 0, 0, None, d.P, 0, B, d.P,
                                                 d.heads_v, d.heads_v, hdv, scale, 0, self.dt, st), "siglip attention (tcgen05)")
                else:
                    cabi.check(L.pg_attention(ptr(att), d.Hv, ptr(qkv), 3 * d.Hv, qkv[:, d.Hv:].data_ptr(),
                                              qkv[:, 2 * d.Hv:].data_ptr(), 3 * d.Hv, d.P * 3 * d.Hv, None, 0, 0,
                                              None, d.P, 0, B, d.P, d.heads_v, d.heads_v, hdv, scale, 0, self.dt, st),
                               "siglip attention")
            with tm("o_proj"):
                self._gemm(h2, att, w["o_w"], w["o_b"], h, cabi.EPI_BIAS_RES)
            with tm("ln"):
                cabi.check(L.pg_layernorm(ptr(ln), ptr(h2), ptr(w["ln2_w"]), ptr(w["ln2_b"]), T, d.Hv, d.eps_v, self.dt, st), "ln2")
            with tm("fc1"):
                self._gemm(mid, ln, w["fc1_w"], w["fc1_b"], None, cabi.EPI_BIAS_GELU)
            with tm("fc2"):
                self._gemm(h, mid, w["fc2_w"], w["fc2_b"], h2, cabi.EPI_BIAS_RES)
        out = self._new(T, d.Hv)
        cabi.check(L.pg_layernorm(ptr(out), ptr(h), ptr(self.v_post_w), ptr(self.v_post_b), T, d.Hv, d.eps_v, self.dt, st), "post_ln")
        return out.view(B, d.P, d.Hv)

    def project(self, feats: torch.Tensor) -> torch.Tensor:
        """PaliGemmaMultiModalProjector.forward (modeling_gemma.py:435-438)."""
        d = self.dims
        x = feats.reshape(-1, d.Hv)
        out = self._new(x.shape[0], self.proj_w.shape[0])
        self._gemm(out, x, self.proj_w, self.proj_b, None, cabi.EPI_BIAS)
        return out.view(*feats.shape[:-1], self.proj_w.shape[0])

    def encode_images(self, pixels: torch.Tensor) -> torch.Tensor:
        """SigLIP tower + projector.  Small batches (the per-request / cache-off case: ~190 launches of 5-15 us) replay
        a CUDA graph captured per batch size, so the launch gaps disappear; large batches are launch-insensitive."""
        B = pixels.shape[0]
        if (not self._vision_graphs_on or B > 8 or self._op_timing or pixels.device != self.device
                or torch.cuda.is_current_stream_capturing()):
            return self.project(self.vision_features(pixels))
        key = (B, tuple(pixels.shape[1:]))
        ent = self._vision_graphs.get(key)
        if ent is None:
            px = torch.empty(pixels.shape, dtype=self.dtype, device=self.device)
            px.copy_(pixels)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):                   # warm-up outside the capture (module loading, smem attributes)
                self.project(self.vision_features(px))
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.project(self.vision_features(px))
            ent = self._vision_graphs[key] = (g, px, out)
        g, px, out = ent
        px.copy_(pixels)
        g.replay()
        return out.clone()

    # ------------------------------------------------------------------ text: q_len >= 1, general
    def text_forward(self, input_ids: torch.Tensor, image_features: Optional[torch.Tensor],
                     kv: Optional[PagedKV], position_value: Optional[int] = None,
                     logits: str = "all") -> torch.Tensor:
        """Embed+merge, L decoder layers, final norm, lm_head for a (B,q) block of new tokens.

        kv=None reproduces `kv_cache=None` (nothing persists).  With an empty cache the new
        tokens get positions 0..q-1; with a non-empty cache every new token gets
        `position_value` (the attention-mask length).  Returns fp32 logits (B,q,V) or, with
        logits='last', (B,1,V)."""
        return self.tp.run(self.text_forward_gen(input_ids, image_features, kv, position_value, logits))

    def text_forward_gen(self, input_ids, image_features, kv, position_value=None, logits="all"):
        """text_forward as a launch generator: yields ("all_reduce", t) / ("all_gather", out, t) where the tensor-parallel
        ranks must meet (TP.run performs them with torch.distributed, LockstepGroup for emulated ranks) and returns
        the logits.  The prefill keeps collectives: its messages are (B*q, D) rows, bandwidth- not latency-sized."""
        d, L, st = self.dims, cabi.lib(), cabi.stream()
        B, q = input_ids.shape
        T = B * q
        ids = input_ids.to(device=self.device, dtype=torch.int64).contiguous()
        temp = kv is None
        if temp:
            kv = self.new_kv(B)
        try:
            if kv.batch != B:
                raise ValueError(f"kv cache holds batch {kv.batch}, input has batch {B}")
            cached = kv.length
            kv.reserve(cached + q)
            if cached == 0:
                pos = torch.arange(q, dtype=torch.int32, device=self.device).repeat(B)
            else:
                pos = torch.full((T,), int(position_value), dtype=torch.int32, device=self.device)
            img = None if image_features is None or image_features.numel() == 0 else \
                image_features.reshape(-1, d.D).to(self.dtype).contiguous()
            x = self._new(T, d.D)
            cabi.check(L.pg_embed_merge(ptr(x), ptr(ids), ptr(self.emb), ptr(img), T, d.D, d.V,
                                        d.image_token_index, d.pad_token_id, 0 if img is None else img.shape[0],
                                        self.img_div, self.normalizer, ptr(self.err_flag), self.dt, st), "embed_merge")
            nq, tp = self.nq_l, self.tp
            res_epi = cabi.EPI_RES if tp.rank == 0 else cabi.EPI_NONE   # the residual enters the sum once
            nqkv = (nq + 2 * d.nkv) * d.hd
            n, qkv, qo = self._new(T, d.D), self._new(T, nqkv), self._new(T, nq * d.hd)
            att, x2, g = self._new(T, nq * d.hd), self._new(T, d.D), self._new(T, self.F_l)
            scale_div = float(math.sqrt(d.hd))
            for li, w in enumerate(self.t_layers):
                cabi.check(L.pg_rmsnorm(ptr(n), ptr(x), ptr(w["ln1"]), T, d.D, d.eps, self.dt, st), "rmsnorm")
                self._gemm(qkv, n, w["qkv"])
                cabi.check(L.pg_rope_append(ptr(qo), ptr(qkv), ptr(self.inv_freq), ptr(pos), ptr(self.k_pool[li]),
                                            ptr(self.v_pool[li]), ptr(kv.page_table), kv.max_pages, self.page_size,
                                            ptr(kv.kv_len), B, q, nq, d.nkv, d.hd, d.max_pos, self.dt, st), "rope_append")
                if self.attn_tc_text and q >= 16:
                    cabi.check(L.pg_attention_tc(ptr(att), nq * d.hd, ptr(qo), T, nq * d.hd, 0, ptr(self.k_pool[li]),
                                                 ptr(self.v_pool[li]), self.num_pages * self.page_size, d.nkv * d.hd, 0, 0,
                                                 d.hd, 0, ptr(kv.page_table), kv.max_pages, self.page_size, ptr(kv.kv_len),
                                                 0, q, B, q, nq, d.nkv, d.hd, scale_div, 1, self.dt, st),
                               "attention (tcgen05)")
                else:
                    cabi.check(L.pg_attention(ptr(att), nq * d.hd, ptr(qo), nq * d.hd, ptr(self.k_pool[li]),
                                              ptr(self.v_pool[li]), 0, 0, ptr(kv.page_table), kv.max_pages, self.page_size,
                                              ptr(kv.kv_len), 0, q, B, q, nq, d.nkv, d.hd, scale_div, 1, self.dt, st),
                               "attention")
                self._gemm(x2, att, w["o"], None, x if tp.rank == 0 else None, res_epi)
                if tp.active:
                    yield ("all_reduce", x2)
                cabi.check(L.pg_rmsnorm(ptr(n), ptr(x2), ptr(w["ln2"]), T, d.D, d.eps, self.dt, st), "rmsnorm")
                self._gemm(g, n, w["gu"], None, None, cabi.EPI_GEGLU)
                self._gemm(x, g, w["down"], None, x2 if tp.rank == 0 else None, res_epi)
                if tp.active:
                    yield ("all_reduce", x)
            kv.kv_len.add_(q)
            kv.length = cached + q
            if logits == "last":
                x = x.view(B, q, d.D)[:, -1].contiguous()
            rows = x.shape[0]
            out = self._new(rows, self.V_l, dtype=torch.float32)
            if rows < self.batched_min:
                # a handful of rows: the weight-streaming lm_head kernel (final norm fused) beats a GEMM tile
                for r0 in range(0, rows, cabi.MAX_DECODE_BATCH):
                    nr = min(cabi.MAX_DECODE_BATCH, rows - r0)
                    cabi.check(L.pg_decode_lmhead(ptr(out[r0:]), ptr(x[r0:]), ptr(self.final_norm), ptr(self.lm_head), nr, d.D,
                                                  self.V_l, d.eps, None, None, None, self.dt, st), "lm_head")
            else:
                h = self._new(rows, d.D)
                cabi.check(L.pg_rmsnorm(ptr(h), ptr(x), ptr(self.final_norm), rows, d.D, d.eps, self.dt, st), "final norm")
                self._gemm(out, h, self.lm_head, out_f32=True)
            if tp.active:  # vocab shards -> full rows
                gathered = self._new(tp.size, rows, self.V_l, dtype=torch.float32)
                yield ("all_gather", gathered, out)
                out = gathered.permute(1, 0, 2).reshape(rows, d.V)
            return out.view(B, -1, d.V)
        finally:
            if temp:
                kv.release()

    def encode_images_dp(self, pixels: torch.Tensor) -> torch.Tensor:
        """Batch-level data parallelism for the vision tower: every rank encodes its contiguous share
        of the images (SigLIP + projector, modeling_gemma.py:568-571) and the (B, P, D) features are all-gathered
        over NCCL, 1 MB per image in bf16.  Fewer images than ranks: every rank encodes them all (replicated)."""
        tp = self.tp
        B = pixels.shape[0]
        if not tp.active or B < tp.size or tp.emulated:
            return self.encode_images(pixels)
        import torch.distributed as dist
        d = self.dims
        lo, hi = shard_batch(B, tp.rank, tp.size)
        per = -(-B // tp.size)                                   # the largest share; smaller ones are padded
        mine = self._new(per, d.P, d.D)
        mine[:hi - lo].copy_(self.encode_images(pixels[lo:hi].contiguous()))
        allf = self._new(tp.size, per, d.P, d.D)
        dist.all_gather_into_tensor(allf.view(tp.size * per, d.P, d.D), mine, group=tp.group)
        if B % tp.size == 0:
            return allf.view(B, d.P, d.D)
        return torch.cat([allf[r, :shard_batch(B, r, tp.size)[1] - shard_batch(B, r, tp.size)[0]] for r in range(tp.size)])

    def check_errors(self, sync: bool = False) -> None:
        """Raise if a kernel flagged an id problem (image token without image rows, id out of range): the reference
        raises from masked_scatter / embedding for the same inputs.  The flag lives in pinned host memory, so the
        check itself never synchronises; callers that are at a synchronisation point anyway (prefill forward, end of
        generate(), the scheduler's per-chunk read-back) pass sync=True to see this call's own kernels, the per-token
        decode path reports at the next call (like the deferred mask check)."""
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        if self._err_np[0] != 0 or (self.fabric is not None and self.fabric.lost_peer()):
            lost = self.fabric is not None and self.fabric.lost_peer()
            self._err_np[0] = 0
            if lost:
                raise RuntimeError("tensor-parallel exchange timed out waiting for a peer rank")
            raise RuntimeError("input_ids held an image token with no image feature left, or an id outside the vocabulary")

    # ------------------------------------------------------------------ decode (q_len == 1)
    def decode_state(self, batch: int) -> "DecodeState":
        key = (batch,)
        if key not in self._decode_states:
            self._decode_states[key] = DecodeState(self, batch)
        return self._decode_states[key]


class DecodeState:
    """Static device buffers + CUDA graph of one decode step for a fixed batch size.

    Everything a step needs that changes between steps (token ids, positions, KV lengths, page
    table, RNG offset) lives in device memory the graph re-reads, so one captured graph serves the
    whole generation and the host never has to synchronise per token."""

    def __init__(self, eng: PaliGemmaEngine, batch: int):
        self.eng, self.B = eng, batch
        d, dev = eng.dims, eng.device
        self.ids = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.pos = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.keys = torch.zeros(batch, dtype=torch.int64, device=dev)  # u64 argmax keys
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.max_hist = 4096
        self.history = torch.zeros((batch, self.max_hist), dtype=torch.int64, device=dev)
        self.logits = torch.zeros((batch, d.V), dtype=torch.float32, device=dev)
        self.probs = None
        self.sampled = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.x = eng._new(batch, d.D)
        self.x2 = eng._new(batch, d.D)
        self.q = eng._new(batch, eng.nq_l * d.hd)
        self.att = eng._new(batch, eng.nq_l * d.hd)
        self.g = eng._new(batch, eng.F_l)
        self.ws = torch.zeros(1, dtype=torch.float32, device=dev)       # (unused by the cluster kernel)
        self.counters = torch.zeros(batch * d.nkv, dtype=torch.int32, device=dev)
        tp = eng.tp
        self.local_logits = self.logits if not tp.active else torch.zeros((batch, eng.V_l), dtype=torch.float32, device=dev)
        self.gather_logits = torch.zeros((tp.size, batch, eng.V_l), dtype=torch.float32, device=dev) if tp.active else None
        self.gather_keys = torch.zeros((tp.size, batch), dtype=torch.int64, device=dev) if tp.active else None
        self.want_full_logits = False  # TP: also all-gather the vocabulary shards of the logits every step (forward() API)
        self.graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self.pf_cap_mb = float(os.environ.get("PG_PF_MB", "0"))   # L2 prefetch distance per launch (0 = off)
        self.kv: Optional[PagedKV] = None
        self.kv_table_ptr = None

    def _pf(self, t: Optional[torch.Tensor], start_mb: float = 0.0, cap_mb: Optional[float] = None) -> None:
        """Hand the next launch a weight region to pull into L2 (pg_set_next_prefetch)."""
        if t is None or self.pf_cap_mb <= 0:
            return
        nbytes = t.numel() * t.element_size()
        off = min(int(start_mb * (1 << 20)), nbytes) & ~127
        n = min(nbytes - off, int((self.pf_cap_mb if cap_mb is None else cap_mb) * (1 << 20))) & ~15
        if n > 0:
            cabi.check(cabi.lib().pg_set_next_prefetch(t.data_ptr() + off, n), "set_next_prefetch")

    # one decode step: kernels only, no host sync, capturable
    def launch_step(self, kv: PagedKV, sample: Optional[tuple] = None, advance: bool = True) -> None:
        self.eng.tp.run(self.step_gen(kv, sample, advance))

    def step_gen(self, kv: PagedKV, sample: Optional[tuple] = None, advance: bool = True):
        """The step as a launch generator (one `yield` per kernel; see dist.TP.run / dist.LockstepGroup)."""
        if self.B >= self.eng.batched_min:
            yield from self._step_batched_gen(kv, sample, advance)
        else:
            yield from self._step_gemv_gen(kv, sample, advance)

    def _fabric(self):
        """The peer-memory exchange when this step can use it (rows fit one slot, one GEMV sub-batch)."""
        fab = self.eng.fabric
        if fab is None or self.B > fab.MAX_ROWS:
            return None
        if self.B < self.eng.batched_min and self.B > cabi.MAX_DECODE_BATCH:
            return None
        return fab

    def layer_gemv_gen(self, li: int, kv: PagedKV, fab, first: bool):
        """Decoder layer `li` of the GEMV step (batch <= 8).  Residual stream: enters in self.x (`first`) or, tensor
        parallel with the exchange, as self.x2 + the pending down_proj partials; leaves in self.x (no exchange) or as
        self.x2 + pending partials (exchange)."""
        eng, d, L, st = self.eng, self.eng.dims, cabi.lib(), cabi.stream()
        B, dt, MB = self.B, self.eng.dt, cabi.MAX_DECODE_BATCH
        tp, nq, F_l = eng.tp, eng.nq_l, eng.F_l
        w = eng.t_layers[li]
        kp, vp = eng.k_pool[li], eng.v_pool[li]
        x, x2 = self.x, self.x2
        stride = 2 * d.L + 2
        scale_div = float(math.sqrt(d.hd))
        cap = self.pf_cap_mb
        nxt = eng.t_layers[li + 1]["qkv"] if li + 1 < len(eng.t_layers) else eng.lm_head
        self._pf(w["o"])                                   # qkv pulls o_proj's weights
        ex_in = fab.x(2 * li, stride) if (fab is not None and not first) else None
        for b0 in range(0, B, MB):
            nb = min(MB, B - b0)
            xin = x2 if ex_in is not None else x
            cabi.check(L.pg_decode_qkv(ptr(self.q[b0:]), ptr(xin[b0:]), ptr(w["ln1"]), ptr(w["qkv"]), ptr(eng.inv_freq),
                                       ptr(self.pos[b0:]), ptr(kp), ptr(vp), ptr(kv.page_table[b0:]), kv.max_pages,
                                       eng.page_size, ptr(kv.kv_len[b0:]), nb, d.D, nq, d.nkv, d.hd, d.eps,
                                       d.max_pos, exref(ex_in), ptr(x[b0:]) if ex_in is not None else None, dt, st),
                       "decode_qkv")
            yield
        self._pf(w["gu"])                                  # attention pulls the head of gate/up
        cabi.check(L.pg_decode_attention(ptr(self.att), ptr(self.q), ptr(kp), ptr(vp), ptr(kv.page_table),
                                         kv.max_pages, eng.page_size, ptr(kv.kv_len), 1, B, nq, d.nkv, d.hd,
                                         scale_div, ptr(self.ws), ptr(self.counters), eng.max_splits, dt, st),
                   "decode_attention")
        yield
        if fab is not None:
            # producer -> consumer pairs: the partial sums travel inside the kernels (csrc/tp_exchange.cuh)
            ex_o, ex_d = fab.x(2 * li + 1, stride), fab.x(2 * li + 2, stride)
            cabi.check(L.pg_gemv_res(None, ptr(self.att), ptr(w["o"]), None, B, d.D, nq * d.hd, exref(ex_o), dt, st), "o_proj")
            yield
        if fab is not None:
            cabi.check(L.pg_decode_gateup(ptr(self.g), ptr(x), ptr(w["ln2"]), ptr(w["gu"]), B, d.D, F_l, d.eps,
                                          exref(ex_o), ptr(x2), dt, st), "gateup")
            yield
            cabi.check(L.pg_gemv_res(None, ptr(self.g), ptr(w["down"]), None, B, d.D, F_l, exref(ex_d), dt, st), "down_proj")
            yield
            return
        for b0 in range(0, B, MB):
            nb = min(MB, B - b0)
            if b0 == 0:
                self._pf(w["gu"], start_mb=cap)            # o_proj pulls the next slice of gate/up
            # tensor parallel over collectives: only rank 0 adds the residual, so it enters the all-reduced sum once
            cabi.check(L.pg_gemv_res(ptr(x2[b0:]), ptr(self.att[b0:]), ptr(w["o"]),
                                     ptr(x[b0:]) if tp.rank == 0 else None, nb, d.D, nq * d.hd, None, dt, st), "o_proj")
            yield
        if tp.active:
            yield ("all_reduce", x2)
        for b0 in range(0, B, MB):
            nb = min(MB, B - b0)
            if b0 == 0:
                self._pf(w["down"], cap_mb=1.5 * cap)      # gate/up pulls the head of down_proj
            cabi.check(L.pg_decode_gateup(ptr(self.g[b0:]), ptr(x2[b0:]), ptr(w["ln2"]), ptr(w["gu"]), nb, d.D,
                                          F_l, d.eps, None, None, dt, st), "gateup")
            yield
            if b0 == 0:
                self._pf(nxt, cap_mb=1.5 * cap)            # down_proj pulls the next layer's qkv / lm_head
            cabi.check(L.pg_gemv_res(ptr(x[b0:]), ptr(self.g[b0:]), ptr(w["down"]),
                                     ptr(x2[b0:]) if tp.rank == 0 else None, nb, d.D, F_l, None, dt, st), "down_proj")
            yield
        if tp.active:
            yield ("all_reduce", x)

    def _step_gemv_gen(self, kv: PagedKV, sample: Optional[tuple], advance: bool):
        eng, d, L, st = self.eng, self.eng.dims, cabi.lib(), cabi.stream()
        B, dt, MB = self.B, self.eng.dt, cabi.MAX_DECODE_BATCH
        tp = eng.tp
        fab = self._fabric()
        stride = 2 * d.L + 2
        if fab is not None:
            cabi.check(L.pg_tp_begin_step(ptr(fab.epoch), st), "tp_begin_step")
            yield
        cabi.check(L.pg_embed_merge(ptr(self.x), ptr(self.ids), ptr(eng.emb), None, B, d.D, d.V,
                                    d.image_token_index, d.pad_token_id, 0, eng.img_div, eng.normalizer,
                                    ptr(eng.err_flag), dt, st), "embed")
        yield
        for li in range(len(eng.t_layers)):
            yield from self.layer_gemv_gen(li, kv, fab, first=li == 0)
        self._pf(eng.t_layers[0]["qkv"])                       # lm_head pulls layer 0 for the next step
        ex_last = fab.x(2 * d.L, stride) if fab is not None else None
        xin = self.x2 if fab is not None else self.x
        for b0 in range(0, B, MB):
            nb = min(MB, B - b0)
            cabi.check(L.pg_decode_lmhead(ptr(self.local_logits[b0:]), ptr(xin[b0:]), ptr(eng.final_norm), ptr(eng.lm_head),
                                          nb, d.D, eng.V_l, d.eps, ptr(self.keys[b0:]), exref(ex_last),
                                          ptr(self.x[b0:]) if fab is not None else None, dt, st), "lm_head")
            yield
        yield from self._finish_gen(kv, sample, advance, fab)

    def _finish_gen(self, kv: PagedKV, sample: Optional[tuple], advance: bool, fab):
        """Token selection + bookkeeping at the end of a step.  self.keys holds this rank's packed (value, local index)
        argmax keys, self.local_logits its vocabulary shard (the whole vocabulary when not tensor parallel)."""
        eng, d, L, st = self.eng, self.eng.dims, cabi.lib(), cabi.stream()
        B, tp = self.B, self.eng.tp
        stride = 2 * d.L + 2
        keys_ex = None
        tp_token = None
        if tp.active:
            if self.want_full_logits or sample is not None:   # vocab shards -> full rows (API parity / sampling)
                yield ("all_gather", self.gather_logits, self.local_logits)
                self.logits.view(B, tp.size, eng.V_l).copy_(self.gather_logits.permute(1, 0, 2))
                yield
            if sample is None:
                if fab is not None and advance:
                    # (max, index) pairs travel through the key exchange; pg_step_advance picks the winner
                    keys_ex = fab.keys(2 * d.L + 1, stride)
                    cabi.check(L.pg_tp_keys_push(ptr(self.keys), B, eng.V_l, exref(keys_ex), st), "tp_keys_push")
                    yield
                else:
                    yield ("all_gather", self.gather_keys, self.keys)
                    tp_token = combine_argmax_keys(self.gather_keys, eng.V_l)
        sampled = None
        if sample is not None:
            temperature, top_p, seed = sample
            if self.probs is None:
                self.probs = torch.empty_like(self.logits)
            cabi.check(L.pg_top_p_sample(ptr(self.sampled), ptr(self.logits), ptr(self.probs), B, d.V,
                                         float(temperature), float(top_p), int(seed), ptr(self.step), None, st),
                       "top_p")
            yield
            sampled = self.sampled
        elif tp_token is not None:
            self.sampled.copy_(tp_token)
            sampled = self.sampled
        if advance:
            cabi.check(L.pg_step_advance(ptr(self.ids), ptr(self.history), self.max_hist, ptr(self.step),
                                         ptr(self.keys), ptr(sampled), ptr(kv.kv_len), ptr(self.pos), B,
                                         exref(keys_ex), st), "step_advance")
            yield

    def _batched_buffers(self):
        eng, d = self.eng, self.eng.dims
        if not hasattr(self, "bn"):
            self.bn = eng._new(self.B, d.D)
            self.bqkv = eng._new(self.B, (eng.nq_l + 2 * d.nkv) * d.hd)
            self.bh = eng._new(self.B, d.D)
            self.part = eng._new(self.B, d.D, dtype=torch.float32) if eng.tp.active else None

    def _reduce_norm(self, fab, out, x_out, x_in, w_norm, index):
        d = self.eng.dims
        ex = fab.x(index, 2 * d.L + 2)
        cabi.check(cabi.lib().pg_rmsnorm_reduce(ptr(out), ptr(x_out), ptr(x_in), ptr(w_norm), self.B, d.D, d.eps, exref(ex),
                                                self.eng.dt, cabi.stream()), "rmsnorm_reduce")

    def _partial_push(self, fab, a, w_mat, index):
        d = self.eng.dims
        self.eng._gemm(self.part, a, w_mat, out_f32=True)
        cabi.check(cabi.lib().pg_tp_push(ptr(self.part), self.B * d.D, exref(fab.x(index, 2 * d.L + 2)), cabi.stream()),
                   "tp_push")

    def layer_batched_gen(self, li: int, kv: PagedKV, fab, first: bool):
        """Decoder layer `li` of the tensor-core step (same stream conventions as layer_gemv_gen)."""
        eng, d, L, st = self.eng, self.eng.dims, cabi.lib(), cabi.stream()
        B, dt = self.B, self.eng.dt
        tp, nq = eng.tp, eng.nq_l
        self._batched_buffers()
        w = eng.t_layers[li]
        scale_div = float(math.sqrt(d.hd))
        x, x2, n = self.x, self.x2, self.bn
        res_epi = cabi.EPI_RES if tp.rank == 0 else cabi.EPI_NONE
        kp, vp = eng.k_pool[li], eng.v_pool[li]
        if fab is not None and not first:
            self._reduce_norm(fab, n, x, x2, w["ln1"], 2 * li)          # x = x2 + sum(down partials of layer li-1)
        else:
            cabi.check(L.pg_rmsnorm(ptr(n), ptr(x), ptr(w["ln1"]), B, d.D, d.eps, dt, st), "rmsnorm")
        yield
        eng._gemm(self.bqkv, n, w["qkv"])
        yield
        cabi.check(L.pg_rope_append(ptr(self.q), ptr(self.bqkv), ptr(eng.inv_freq), ptr(self.pos), ptr(kp), ptr(vp),
                                    ptr(kv.page_table), kv.max_pages, eng.page_size, ptr(kv.kv_len), B, 1, nq, d.nkv,
                                    d.hd, d.max_pos, dt, st), "rope_append")
        yield
        cabi.check(L.pg_decode_attention(ptr(self.att), ptr(self.q), ptr(kp), ptr(vp), ptr(kv.page_table),
                                         kv.max_pages, eng.page_size, ptr(kv.kv_len), 1, B, nq, d.nkv, d.hd,
                                         scale_div, ptr(self.ws), ptr(self.counters), eng.max_splits, dt, st),
                   "decode_attention")
        yield
        if fab is not None:
            self._partial_push(fab, self.att, w["o"], 2 * li + 1)
            yield
            self._reduce_norm(fab, n, x2, x, w["ln2"], 2 * li + 1)      # x2 = x + sum(o_proj partials)
            yield
        else:
            eng._gemm(x2, self.att, w["o"], None, x if tp.rank == 0 else None, res_epi)
            yield
            if tp.active:
                yield ("all_reduce", x2)
            cabi.check(L.pg_rmsnorm(ptr(n), ptr(x2), ptr(w["ln2"]), B, d.D, d.eps, dt, st), "rmsnorm")
            yield
        eng._gemm(self.g, n, w["gu"], None, None, cabi.EPI_GEGLU)
        yield
        if fab is not None:
            self._partial_push(fab, self.g, w["down"], 2 * li + 2)
            yield
        else:
            eng._gemm(x, self.g, w["down"], None, x2 if tp.rank == 0 else None, res_epi)
            yield
            if tp.active:
                yield ("all_reduce", x)

    def _step_batched_gen(self, kv: PagedKV, sample: Optional[tuple], advance: bool):
        """Decode step for batches above the GEMV tile (BASELINE configs[3]: batch 32): the projections run as
        skinny GEMMs (tcgen05 for 16-bit dtypes: weights stream through TMA once for the whole batch), attention
        stays the cluster kernel.  Same rounding points as the GEMV path; capturable (static buffers).
        Tensor parallel: the o_proj / down_proj GEMMs leave fp32 partials in local memory, pg_tp_push hands them to
        every rank and pg_rmsnorm_reduce (sum + residual + RMSNorm) replaces the all-reduce and the norm launch."""
        eng, d, L, st = self.eng, self.eng.dims, cabi.lib(), cabi.stream()
        B, dt = self.B, self.eng.dt
        tp, nq, F_l = eng.tp, eng.nq_l, eng.F_l
        fab = self._fabric()
        stride = 2 * d.L + 2
        self._batched_buffers()
        if fab is not None:
            cabi.check(L.pg_tp_begin_step(ptr(fab.epoch), st), "tp_begin_step")
            yield
        cabi.check(L.pg_embed_merge(ptr(self.x), ptr(self.ids), ptr(eng.emb), None, B, d.D, d.V,
                                    d.image_token_index, d.pad_token_id, 0, eng.img_div, eng.normalizer,
                                    ptr(eng.err_flag), dt, st), "embed")
        yield
        for li in range(len(eng.t_layers)):
            yield from self.layer_batched_gen(li, kv, fab, first=li == 0)
        x, x2 = self.x, self.x2
        if fab is not None:
            self._reduce_norm(fab, self.bh, x, x2, eng.final_norm, 2 * d.L)
        else:
            cabi.check(L.pg_rmsnorm(ptr(self.bh), ptr(x), ptr(eng.final_norm), B, d.D, d.eps, dt, st), "final norm")
        yield
        eng._gemm(self.local_logits, self.bh, eng.lm_head, out_f32=True)
        yield
        if sample is None:
            # packed (value, index) keys of this rank's logits; not tensor parallel: also the token itself
            cabi.check(L.pg_argmax(None if tp.active else ptr(self.sampled), ptr(self.local_logits), ptr(self.keys), B,
                                   eng.V_l, st), "argmax")
            yield
        if not tp.active and sample is None:
            if advance:
                cabi.check(L.pg_step_advance(ptr(self.ids), ptr(self.history), self.max_hist, ptr(self.step),
                                             ptr(self.keys), ptr(self.sampled), ptr(kv.kv_len), ptr(self.pos), B,
                                             None, st), "step_advance")
                yield
            return
        yield from self._finish_gen(kv, sample, advance, fab)

    def bind(self, kv: PagedKV, next_ids: torch.Tensor, position: int) -> None:
        """Point the step at a cache and seed ids / positions (host -> device, outside the graph)."""
        self.kv = kv
        self.ids.copy_(next_ids.reshape(-1).to(torch.int64))
        self.pos.fill_(int(position))
        self.step.zero_()
        self.keys.zero_()

    def run_steps(self, kv: PagedKV, n_steps: int, sample: Optional[tuple] = None, use_graph: bool = True) -> None:
        """Advance n_steps tokens entirely on the device (tokens land in self.history)."""
        if n_steps > self.max_hist:
            raise ValueError("too many steps for the history buffer")
        kv.reserve(kv.length + n_steps)
        if not use_graph:
            for _ in range(n_steps):
                self.launch_step(kv, sample)
        else:
            key = (kv.page_table.data_ptr(), kv.kv_len.data_ptr(), kv.max_pages, sample, self.want_full_logits)
            g = self.graphs.get(key)
            if g is None:
                # warm up on a side stream (lazy module loading, smem attributes), undo its effects
                saved = [t.clone() for t in (self.ids, self.pos, self.step, self.keys, kv.kv_len, self.history[:, :1])]
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self.launch_step(kv, sample)
                torch.cuda.current_stream().wait_stream(s)
                for t, v in zip((self.ids, self.pos, self.step, self.keys, kv.kv_len, self.history[:, :1]), saved):
                    t.copy_(v)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.launch_step(kv, sample)
                for t, v in zip((self.ids, self.pos, self.step, self.keys, kv.kv_len, self.history[:, :1]), saved):
                    t.copy_(v)
                if len(self.graphs) > 8:
                    self.graphs.clear()
                self.graphs[key] = g
            for _ in range(n_steps):
                g.replay()
        kv.length += n_steps
