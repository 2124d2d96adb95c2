"""Tensor-parallel plumbing for the Gemma decoder (the reference itself is single-device).

Sharding (tp = world size, one process per GPU):
  q_proj rows by query head, o_proj columns by query head   -> sum over ranks of (B*q, D) after o_proj
  gate/up rows, down_proj columns by intermediate feature    -> sum over ranks after down_proj
  lm_head rows by vocabulary                                 -> exchange of (max, index) pairs / all-gather of logits
  k_proj / v_proj and the KV cache are replicated (one KV head); embeddings and the vision tower too
  (vision batches are data-parallel: `PaliGemmaEngine.encode_images_dp`).

Decode steps do NOT call a collective for the 36 sums per token: the GEMV epilogue stores its fp32 partial straight
into every rank's exchange buffer over NVLink and the next kernel's RMSNorm prologue sums what it finds in local memory
(csrc/tp_exchange.cuh, `Fabric` below).  torch.distributed / NCCL carries the rest: the prefill's large all-reduces, the
optional all-gather of full logits, and the rendezvous that maps the peers' buffers.

`LockstepGroup` runs N ranks on ONE GPU for tests: the engines expose their launch sequences as generators that yield
after every kernel, and the group advances all ranks one kernel at a time on a single stream, so every rank's producer
has run before any rank's consumer looks for its data (kernels that wait for each other must never be launched
separately on one GPU otherwise).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch


class Exchange(C.Structure):
    """`pg_tp_exchange` of include/pg_b200.h."""
    _fields_ = [("peers", C.c_void_p), ("epoch", C.c_void_p), ("err_dev", C.c_void_p), ("err_host", C.c_void_p),
                ("region_off", C.c_longlong), ("slot_bytes", C.c_longlong),
                ("rank", C.c_int), ("size", C.c_int), ("index", C.c_int), ("stride", C.c_int)]


class Fabric:
    """One rank's exchange buffer + the table of every rank's buffer address.

    Layout: region "x"    [2 parities][size ranks][MAX_ROWS * D words of 8 bytes]   partial rows of o_proj / down_proj
            region "keys" [2 parities][size ranks][MAX_ROWS * 2 words]               packed (value, index) argmax keys"""
    MAX_ROWS = 64

    def __init__(self, rank: int, size: int, D: int, buf: torch.Tensor, peers_dev: int, keepalive=None):
        if not 2 <= size <= 8:
            raise ValueError("the peer-memory exchange supports 2..8 ranks")
        self.rank, self.size, self.D = rank, size, D
        self.buf, self.peers_dev, self._keepalive = buf, peers_dev, keepalive
        dev = buf.device
        self.x_slot = self.MAX_ROWS * D * 8
        self.keys_slot = self.MAX_ROWS * 16
        self.keys_off = 2 * size * self.x_slot
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.err_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.err_host = torch.zeros(1, dtype=torch.int32)
        if dev.type == "cuda":
            self.err_host = self.err_host.pin_memory()      # device-mapped: the kernels' timeout flag is readable without a sync
        self._err_np = self.err_host.numpy()
        self._cache: Dict[tuple, Exchange] = {}

    @staticmethod
    def nbytes(size: int, D: int) -> int:
        return 2 * size * (Fabric.MAX_ROWS * D * 8 + Fabric.MAX_ROWS * 16)

    def _ex(self, off: int, slot: int, index: int, stride: int) -> Exchange:
        key = (off, index, stride)
        e = self._cache.get(key)
        if e is None:
            e = self._cache[key] = Exchange(self.peers_dev, self.epoch.data_ptr(), self.err_dev.data_ptr(),
                                            self.err_host.data_ptr(), off, slot, self.rank, self.size, index, stride)
        return e

    def x(self, index: int, stride: int) -> Exchange:
        return self._ex(0, self.x_slot, index, stride)

    def keys(self, index: int, stride: int) -> Exchange:
        return self._ex(self.keys_off, self.keys_slot, index, stride)

    def lost_peer(self) -> bool:
        return bool(self._err_np[0] != 0)

    # ---- construction
    @staticmethod
    def symmetric(rank: int, size: int, D: int, device) -> "Fabric":
        """Real ranks, one per GPU: torch's symmetric-memory rendezvous maps every peer's buffer into this process
        (plumbing only; the kernels that use the mapping are ours)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        buf = symm.empty(Fabric.nbytes(size, D), dtype=torch.uint8, device=device)
        buf.zero_()
        hdl = symm.rendezvous(buf, dist.group.WORLD)
        fab = Fabric(rank, size, D, buf, int(hdl.buffer_ptrs_dev), keepalive=hdl)
        torch.cuda.synchronize(device)
        dist.barrier()      # every rank's flags are zero before anyone produces
        return fab

    @staticmethod
    def emulated(size: int, D: int, device) -> List["Fabric"]:
        """N ranks on one GPU (tests): plain device buffers, driven by LockstepGroup."""
        bufs = [torch.zeros(Fabric.nbytes(size, D), dtype=torch.uint8, device=device) for _ in range(size)]
        table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=device)
        return [Fabric(r, size, D, bufs[r], table.data_ptr(), keepalive=(bufs, table)) for r in range(size)]


class TP:
    """Rank / size of the tensor-parallel group and its two transports: `fabric` (peer-memory exchange used by the
    decode kernels; None = every sum goes through all_reduce) and torch.distributed (all_reduce / all_gather)."""

    def __init__(self, rank: int = 0, size: int = 1, group=None, fabric: Optional[Fabric] = None, emulated: bool = False):
        self.rank, self.size, self.group = rank, size, group
        self.fabric = fabric
        self.emulated = emulated          # collectives are performed by a LockstepGroup, not torch.distributed

    @property
    def active(self) -> bool:
        return self.size > 1

    def make_fabric(self, D: int, device) -> Optional[Fabric]:
        """The exchange of ONE engine (its own buffers and step counter: sequence numbers of two models never mix).
        Collective: every rank builds its engines in the same order.  PG_TP_EXCHANGE=nccl keeps the collective path
        (A/B runs); emulated ranks get their fabric from the test (Fabric.emulated)."""
        import os
        if not self.active:
            return None
        if self.emulated:
            return self.fabric
        if os.environ.get("PG_TP_EXCHANGE", "peer") == "nccl" or self.size > 8:
            return None
        return Fabric.symmetric(self.rank, self.size, D, device)

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.size > 1:
            if self.emulated:
                raise RuntimeError("emulated ranks: drive the engine generators through LockstepGroup")
            import torch.distributed as dist
            dist.all_reduce(t, group=self.group)
        return t

    def all_gather(self, out: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """out: [size, *t.shape] contiguous."""
        if self.size > 1:
            if self.emulated:
                raise RuntimeError("emulated ranks: drive the engine generators through LockstepGroup")
            import torch.distributed as dist
            t = t.contiguous()
            flat = out.view(self.size * t.shape[0], *t.shape[1:]) if t.dim() > 0 else out
            dist.all_gather_into_tensor(flat, t, group=self.group)  # concatenation along dim 0
        else:
            out.copy_(t.unsqueeze(0))
        return out

    def run(self, gen):
        """Drive one rank's launch generator with the real collectives; returns the generator's return value."""
        try:
            op = next(gen)
            while True:
                if op is not None:
                    if op[0] == "all_reduce":
                        self.all_reduce(op[1])
                    elif op[0] == "all_gather":
                        self.all_gather(op[1], op[2])
                    else:
                        raise RuntimeError(f"unknown collective {op[0]!r}")
                op = next(gen)
        except StopIteration as stop:
            return stop.value


class LockstepGroup:
    """All ranks of a tensor-parallel group in ONE process on ONE GPU (tests).  `run(gens)` advances the ranks' launch
    generators one kernel at a time, in rank order, on the current stream; collectives the generators ask for are
    performed here (sum in rank order / concatenation), so no torch.distributed is needed."""

    def run(self, gens: list) -> list:
        n = len(gens)
        results, done = [None] * n, [False] * n
        while True:
            ops = []
            for r, g in enumerate(gens):
                if done[r]:
                    ops.append(None)
                    continue
                try:
                    ops.append(next(g))
                except StopIteration as stop:
                    results[r], done[r] = stop.value, True
                    ops.append(None)
            if all(done):
                return results
            if any(done):
                raise RuntimeError("ranks issued different launch sequences")
            kinds = {None if o is None else o[0] for o in ops}
            if len(kinds) != 1:
                raise RuntimeError(f"ranks diverged: {kinds}")
            kind = kinds.pop()
            if kind == "all_reduce":
                total = ops[0][1].clone()
                for o in ops[1:]:
                    total += o[1]
                for o in ops:
                    o[1].copy_(total)
            elif kind == "all_gather":
                for o in ops:
                    for r, src in enumerate(ops):
                        o[1][r].copy_(src[2])
            elif kind is not None:
                raise RuntimeError(f"unknown collective {kind!r}")


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


def shard_batch(n: int, rank: int, size: int):
    """Contiguous share [lo, hi) of n independent items (images of a vision batch) for `rank`; the first n % size
    ranks take one extra item."""
    base, extra = divmod(n, size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def combine_argmax_keys(keys_all: torch.Tensor, v_local: int) -> torch.Tensor:
    """keys_all: int64 [size, B] holding the packed u64 (ordered value bits << 32 | ~local index).
    Returns the global argmax token per batch row; equal values go to the lowest global index, i.e.
    torch.argmax's tie rule on the gathered logits.  (Host-side statement of what pg_tp_keys_push +
    pg_step_advance do on the device; used by tests and the NCCL fallback path.)"""
    size = keys_all.shape[0]
    val = (keys_all >> 32) & 0xFFFFFFFF                      # ordered value bits, 0..2^32-1
    idx = 0xFFFFFFFF - (keys_all & 0xFFFFFFFF)               # local index
    ranks = torch.arange(size, device=keys_all.device, dtype=torch.int64)[:, None]
    gidx = idx + ranks * v_local
    # larger value first, then smaller global index: one comparable integer (val < 2^32, gidx < 2^31)
    score = val * (1 << 31) + ((1 << 31) - 1 - gidx)
    best = score.argmax(dim=0, keepdim=True)
    return gidx.gather(0, best).squeeze(0)
