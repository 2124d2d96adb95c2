"""Continuous batching around the decode graph.

The reference serves one request at a time (`inference.py:34-85`: batch 1, one image, the processor asserts a
single prompt, `processing_paligemma.py:80`).  Here requests of different prompt lengths share ONE captured decode
step: a fixed number of slots, each slot a row of a static page table over the engine's paged KV pool.

  * admission: queued requests with the SAME prompt length are prefilled TOGETHER, as many as there are free slots
    (one vision-tower batch, one text_forward over (B, N) tokens: the weights stream once for all of them; rows of a
    batch never mix, so every row equals its own batch-1 prefill) into pages of their own; each page list is then moved
    into a free slot's row, the slot's KV length / next id / position are written on the device and the sequence
    joins the running batch.  Prompts of different lengths are prefilled in separate calls: padding them into one
    call would need a per-sequence attention mask, which the reference rules out (`modeling_gemma.py:559`, and its
    attention is unmasked: padded positions would be attended);
  * every iteration replays the decode graph `chunk` times for all slots (idle slots compute on a parking page and
    are reset between chunks), reads the chunk's tokens back once, retires sequences that hit EOS or their token
    budget and frees their pages;
  * per-sequence results are identical to `generate()` run on that request alone: rows of a decode step never mix.

Greedy or top-p (one setting for the whole batch: the sampling kernel takes a single temperature / top_p).
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from typing import Deque, Dict, List, Optional

import torch

from .engine import PagedKV, PaliGemmaEngine


class SlotKV(PagedKV):
    """PagedKV with a STATIC page table (fixed pointer and shape, so the decode graph is captured once) whose rows
    are (re)assigned to sequences.  Row `s` holds the pages of the sequence in slot s, padded with the parking page."""

    def __init__(self, engine: PaliGemmaEngine, slots: int, max_tokens: int):
        super().__init__(engine, slots)
        self.max_pages = (max_tokens + engine.page_size - 1) // engine.page_size
        self.parking = engine._alloc_pages(1)[0]
        self.page_table = torch.full((slots, self.max_pages), self.parking, dtype=torch.int32, device=engine.device)
        self.host_len = [0] * slots          # KV entries per slot (host mirror; 0 = idle)
        self.active = [False] * slots

    def reserve(self, total_len: int) -> None:   # DecodeState.run_steps calls this with the uniform-length view
        return

    def assign(self, slot: int, pages: List[int], length: int) -> None:
        if len(pages) > self.max_pages:
            raise RuntimeError(f"sequence needs {len(pages)} pages, a slot holds {self.max_pages}")
        self.pages[slot] = list(pages)
        row = pages + [self.parking] * (self.max_pages - len(pages))
        self.page_table[slot].copy_(torch.tensor(row, dtype=torch.int32), non_blocking=False)
        self.kv_len[slot:slot + 1].fill_(length)
        self.host_len[slot] = length
        self.active[slot] = True

    def ensure(self, slot: int, total_len: int) -> None:
        """Pages for `total_len` entries of the sequence in `slot` (called before the steps that need them)."""
        eng = self.engine
        need = (total_len + eng.page_size - 1) // eng.page_size
        have = len(self.pages[slot])
        if need > self.max_pages:
            raise RuntimeError(f"sequence of {total_len} tokens exceeds the slot capacity")
        if need > have:
            new = eng._alloc_pages(need - have)
            self.pages[slot].extend(new)
            self.page_table[slot, have:need].copy_(torch.tensor(new, dtype=torch.int32))

    def retire(self, slot: int) -> None:
        self.engine._free_pages(self.pages[slot])
        self.pages[slot] = []
        self.page_table[slot].fill_(self.parking)
        self.kv_len[slot:slot + 1].zero_()
        self.host_len[slot] = 0
        self.active[slot] = False

    def release(self) -> None:
        if getattr(self, "parking", None) is not None and self.engine is not None:
            self.engine._free_pages([self.parking])
            self.parking = None
        super().release()


@dataclass
class Request:
    input_ids: torch.Tensor                 # (1, N) int64
    pixel_values: Optional[torch.Tensor]    # (1, 3, S, S) or None
    max_new_tokens: int
    eos_token_id: Optional[int] = None
    rid: int = -1
    tokens: List[int] = field(default_factory=list)
    done: bool = False


class ContinuousBatcher:
    """`submit()` requests, `run()` until all are finished; `finished[rid].tokens` holds the generated ids
    (the EOS token included when one stops the sequence, as inference.py:70-72 appends before breaking)."""

    def __init__(self, engine: PaliGemmaEngine, slots: int = 8, max_tokens: int = 2048, chunk: int = 8,
                 do_sample: bool = False, temperature: float = 0.8, top_p: float = 0.9, seed: int = 0):
        self.eng = engine
        self.slots = slots
        self.chunk = max(1, chunk)
        self.sample = (float(temperature), float(top_p), int(seed)) if do_sample else None
        self.kv = SlotKV(engine, slots, max_tokens)
        self.ds = engine.decode_state(slots)
        self.ds.kv = self.kv
        self.ds.ids.fill_(1)
        self.ds.pos.fill_(1)
        self.ds.step.zero_()      # RNG offset + ring index of the history buffer: monotonic, never reset
        self.ds.keys.zero_()
        self._hist_pos = 0        # host mirror of ds.step
        self.queue: Deque[Request] = deque()
        self.running: Dict[int, Request] = {}     # slot -> request
        self.finished: Dict[int, Request] = {}
        self._next_rid = 0
        self.steps_run = 0
        self.prefill_calls = 0    # text_forward calls spent on admission (batched over equal prompt lengths)

    def submit(self, input_ids: torch.Tensor, pixel_values: Optional[torch.Tensor], max_new_tokens: int,
               eos_token_id: Optional[int] = None) -> int:
        if input_ids.dim() != 2 or input_ids.shape[0] != 1:
            raise ValueError("one prompt per request: input_ids must be (1, N)")
        if max_new_tokens < 1:
            raise ValueError("max_new_tokens must be >= 1")
        cap = self.kv.max_pages * self.eng.page_size
        if input_ids.shape[1] + int(max_new_tokens) > cap:
            raise ValueError(f"prompt ({input_ids.shape[1]}) + max_new_tokens ({max_new_tokens}) exceeds the slot "
                             f"capacity of {cap} tokens (ContinuousBatcher(max_tokens=...))")
        r = Request(input_ids, pixel_values, int(max_new_tokens), eos_token_id, self._next_rid)
        self._next_rid += 1
        self.queue.append(r)
        return r.rid

    # ---- admission: prefill (batched over equal prompt lengths), then move the pages into slots
    @torch.no_grad()
    def _admit(self, slots: List[int], reqs: List[Request]) -> None:
        """Prefill `reqs` (all of one prompt length, all with or all without an image) in ONE forward and seat them
        in `slots`.  A request that is finished by its first token (EOS / budget 1) never takes a slot."""
        eng = self.eng
        from .generate import _pick
        B = len(reqs)
        ids = torch.cat([r.input_ids.to(eng.device) for r in reqs], 0)
        N = ids.shape[1]
        kvb = eng.new_kv(B)
        try:
            kvb.reserve(N + 1)
            feats = None
            if reqs[0].pixel_values is not None:
                feats = eng.encode_images(torch.cat([r.pixel_values.to(eng.device) for r in reqs], 0))
            logits = eng.text_forward(ids, feats, kvb, logits="last")
            # the first token's draw takes its Philox offset from the request id (never the same uniform twice)
            first = _pick(eng, logits[:, -1, :].contiguous(), self.sample, step=((reqs[0].rid + 1) << 16) & 0x3fffffff)
            pages, kvb.pages = kvb.pages, [[] for _ in range(B)]      # ownership moves to the slots
        except Exception:
            for r in reversed(reqs):
                self.queue.appendleft(r)                               # admission failed (e.g. pool exhausted): nothing is lost
            raise
        finally:
            kvb.release()
        toks = first.tolist()
        free = list(slots)
        self.prefill_calls += 1
        for b, r in enumerate(reqs):
            tok = int(toks[b])
            r.tokens.append(tok)
            if (r.eos_token_id is not None and tok == r.eos_token_id) or r.max_new_tokens == 1:
                eng._free_pages(pages[b])
                r.done = True
                self.finished[r.rid] = r
                continue
            slot = free.pop(0)
            self.kv.assign(slot, pages[b], N)
            self.ds.ids[slot:slot + 1].copy_(first[b:b + 1])
            self.ds.pos[slot:slot + 1].fill_(N + 1)        # after t tokens the next one is fed at position N + t (Q3)
            self.running[slot] = r

    def _next_group(self, n_free: int) -> List[Request]:
        """Up to n_free queued requests sharing the head request's prompt length and image-ness, in queue order."""
        head = self.queue[0]
        key = (head.input_ids.shape[1], head.pixel_values is None)
        group, rest = [], deque()
        while self.queue:
            r = self.queue.popleft()
            if len(group) < n_free and (r.input_ids.shape[1], r.pixel_values is None) == key:
                group.append(r)
            else:
                rest.append(r)
        self.queue = rest
        return group

    @torch.no_grad()
    def step(self) -> None:
        """One scheduler iteration: admit into free slots, run one chunk of decode steps, collect, retire."""
        while self.queue:
            free = [s for s in range(self.slots) if s not in self.running]
            if not free:
                break
            self._admit(free, self._next_group(len(free)))
        if not self.running:
            return
        n = min(self.chunk, min(r.max_new_tokens - len(r.tokens) for r in self.running.values()))
        for slot in self.running:
            self.kv.ensure(slot, self.kv.host_len[slot] + n)
        # idle slots: park them at length 0 / position 1 so their rows never walk off the page table
        idle = [s for s in range(self.slots) if s not in self.running]
        if idle:
            idx = torch.tensor(idle, dtype=torch.int64, device=self.eng.device)
            self.kv.kv_len.index_fill_(0, idx, 0)
            self.ds.pos.index_fill_(0, idx, 1)
            self.ds.ids.index_fill_(0, idx, 1)
        self.ds.run_steps(self.kv, n, sample=self.sample)
        self.steps_run += n
        # ring buffer; the device counter is a 32-bit integer that wraps, so its host mirror wraps with it
        cols = ((torch.arange(n, device=self.eng.device) + self._hist_pos) & 0xFFFFFFFF) % self.ds.max_hist
        hist = self.ds.history[:, cols].cpu()               # one read-back per chunk
        self.eng.check_errors()                             # the read-back synchronised: bad ids raise here
        self._hist_pos = (self._hist_pos + n) & 0xFFFFFFFF
        for slot, r in list(self.running.items()):
            self.kv.host_len[slot] += n
            for t in hist[slot].tolist():
                r.tokens.append(int(t))
                if (r.eos_token_id is not None and t == r.eos_token_id) or len(r.tokens) >= r.max_new_tokens:
                    r.done = True
                    break
            if r.done:
                self.kv.retire(slot)
                del self.running[slot]
                self.finished[r.rid] = r

    def run(self) -> Dict[int, Request]:
        while self.queue or self.running:
            self.step()
        return self.finished

    def close(self) -> None:
        for slot in list(self.running):
            self.kv.retire(slot)
        self.running.clear()
        self.kv.release()
