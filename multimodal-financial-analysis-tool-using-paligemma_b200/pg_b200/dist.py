"""Tensor-parallel plumbing for the Gemma decoder (the reference itself is single-device).

Sharding (tp = world size, one process per GPU, torch.distributed / NCCL over NVLink):
  q_proj rows by query head, o_proj columns by query head   -> all-reduce(sum) of (B*q, D) after o_proj
  gate/up rows, down_proj columns by intermediate feature    -> all-reduce(sum) after down_proj
  lm_head rows by vocabulary                                 -> all-gather of logits / (max, index) pairs
  k_proj / v_proj and the KV cache are replicated (one KV head); embeddings and the vision tower too.
The residual is added on rank 0 only, before the all-reduce, so it enters the sum exactly once.
"""
from __future__ import annotations

from typing import Optional

import torch


class OneShot:
    """Symmetric (peer-mapped) buffers + step counter for pg_allreduce_oneshot.  torch's symmetric-memory
    rendezvous is only the plumbing that exchanges the peer mappings; the all-reduce kernel is ours."""
    CAP = 1 << 16          # bytes per data slot: (B<=8, 2048) bf16/fp32 partials fit

    def __init__(self, rank: int, size: int, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.rank, self.size = rank, size
        self.buf = symm.empty(2 * self.CAP + 2 * 16 * 4, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.peers_dev = int(self.hdl.buffer_ptrs_dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier()      # every rank's flags are zero before anyone signals

    def fits(self, t: torch.Tensor) -> bool:
        return t.is_contiguous() and t.numel() * t.element_size() <= self.CAP and t.numel() % 8 == 0


class TP:
    def __init__(self, rank: int = 0, size: int = 1, group=None):
        self.rank, self.size, self.group = rank, size, group
        self.oneshot = None
        self._oneshot_tried = False

    def _maybe_oneshot(self, device):
        """PG_TP_ALLREDUCE=oneshot selects the peer-memory kernel (parity-green on 2 GPUs, err flag 0); the
        default stays NCCL: as a separate launch the one-shot kernel measured 1.166 ms/step vs 1.116 ms
        for NCCL at tp=2 — it only pays once it is fused into the GEMV epilogue (next round)."""
        import os
        if self._oneshot_tried:
            return self.oneshot
        self._oneshot_tried = True
        if os.environ.get("PG_TP_ALLREDUCE", "nccl") != "oneshot" or self.size > 16:
            return None
        try:
            self.oneshot = OneShot(self.rank, self.size, device)
        except Exception as e:  # noqa: BLE001  (no peer access / unsupported allocator: NCCL still works)
            import warnings
            warnings.warn(f"one-shot all-reduce unavailable ({e!r}); using NCCL")
            self.oneshot = None
        return self.oneshot

    @property
    def active(self) -> bool:
        return self.size > 1

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.size > 1:
            one = self._maybe_oneshot(t.device) if t.is_cuda else None
            if one is not None and one.fits(t):
                from . import _cabi as cabi
                cabi.check(cabi.lib().pg_allreduce_oneshot(t.data_ptr(), one.peers_dev, self.rank, self.size, t.numel(),
                                                           one.CAP, one.step.data_ptr(), one.err.data_ptr(),
                                                           cabi.DTYPE_CODE[t.dtype], cabi.stream()), "allreduce_oneshot")
                return t
            import torch.distributed as dist
            dist.all_reduce(t, group=self.group)
        return t

    def all_gather(self, out: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """out: [size, *t.shape] contiguous."""
        if self.size > 1:
            import torch.distributed as dist
            t = t.contiguous()
            flat = out.view(self.size * t.shape[0], *t.shape[1:]) if t.dim() > 0 else out
            dist.all_gather_into_tensor(flat, t, group=self.group)  # concatenation along dim 0
        else:
            out.copy_(t.unsqueeze(0))
        return out


def check_divisible(dims, size: int) -> None:
    for name, v in (("num_attention_heads", dims.nq), ("intermediate_size", dims.F), ("vocab_size", dims.V)):
        if v % size:
            raise ValueError(f"tensor parallel size {size} does not divide {name}={v}")
    if (dims.V // size) % 2:
        raise ValueError("vocabulary shard must be even")


def shard_rows(w: torch.Tensor, rank: int, size: int) -> torch.Tensor:
    n = w.shape[0] // size
    return w[rank * n:(rank + 1) * n]


def shard_cols(w: torch.Tensor, rank: int, size: int) -> torch.Tensor:
    n = w.shape[1] // size
    return w[:, rank * n:(rank + 1) * n].contiguous()


def shard_text_layer(q, k, v, o, gate, up, down, rank: int, size: int):
    """One decoder layer's matrices -> this rank's (qkv, o, gate_up, down)."""
    qkv = torch.cat([shard_rows(q, rank, size), k, v], 0)
    gu = torch.cat([shard_rows(gate, rank, size), shard_rows(up, rank, size)], 0)
    return qkv, shard_cols(o, rank, size), gu, shard_cols(down, rank, size)


def combine_argmax_keys(keys_all: torch.Tensor, v_local: int) -> torch.Tensor:
    """keys_all: int64 [size, B] holding the packed u64 (ordered value bits << 32 | ~local index).
    Returns the global argmax token per batch row; equal values go to the lowest global index, i.e.
    torch.argmax's tie rule on the gathered logits."""
    size = keys_all.shape[0]
    val = (keys_all >> 32) & 0xFFFFFFFF                      # ordered value bits, 0..2^32-1
    idx = 0xFFFFFFFF - (keys_all & 0xFFFFFFFF)               # local index
    ranks = torch.arange(size, device=keys_all.device, dtype=torch.int64)[:, None]
    gidx = idx + ranks * v_local
    # larger value first, then smaller global index: one comparable integer (val < 2^32, gidx < 2^31)
    score = val * (1 << 31) + ((1 << 31) - 1 - gidx)
    best = score.argmax(dim=0, keepdim=True)
    return gidx.gather(0, best).squeeze(0)
