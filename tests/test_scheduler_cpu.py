"""Host logic of the continuous-batching scheduler on the CPU: a stub engine whose "model" emits a deterministic token
stream per request (token t of a request = f(prompt, t)) stands in for the CUDA engine, so admission, chunked stepping,
EOS / budget retirement, slot reuse, the static page table and the page accounting are exercised without a GPU."""
import torch

from pg_b200 import scheduler as S


def _stream(prompt_sum: int, t: int) -> int:
    return 3 + (prompt_sum * 7 + t * 13) % 50


class StubDecodeState:
    """Mimics DecodeState: ids / pos / step / keys / history buffers and run_steps(kv, n) advancing every slot."""
    def __init__(self, eng, batch):
        self.eng, self.B, self.max_hist = eng, batch, 64
        self.ids = torch.zeros(batch, dtype=torch.int64)
        self.pos = torch.zeros(batch, dtype=torch.int32)
        self.step = torch.zeros(1, dtype=torch.int32)
        self.keys = torch.zeros(batch, dtype=torch.int64)
        self.history = torch.zeros((batch, self.max_hist), dtype=torch.int64)
        self.kv = None
        self.calls = []

    def run_steps(self, kv, n, sample=None):
        self.calls.append(n)
        for _ in range(n):
            col = int(self.step.item()) % self.max_hist      # ring buffer, like step_advance_kernel
            for b in range(self.B):
                L, pos = int(kv.kv_len[b]), int(self.pos[b])
                # an active row must own a real (non-parking) page for the entry it is about to write
                if kv.host_len[b] > 0:
                    page = int(kv.page_table[b, L // self.eng.page_size])
                    assert page != kv.parking and page in kv.pages[b], (b, L, page)
                    assert pos == L + 1                      # Q3: position = mask length
                owner = self.eng.slot_owner.get(b)
                tok = _stream(owner[0], pos - owner[1] - 1 + 1) if owner and kv.host_len[b] > 0 else 1
                self.history[b, col] = tok
                self.ids[b] = tok
            kv.kv_len += 1
            self.pos += 1
            self.step += 1


class StubEngine:
    page_size = 4
    device = torch.device("cpu")

    def __init__(self, pages=64):
        self._free = list(range(pages - 1, -1, -1))
        self.slot_owner = {}
        self._ds = {}
        self.prefills = []
        self.prefill_batches = []

    def _alloc_pages(self, n):
        assert n <= len(self._free), "pool exhausted"
        return [self._free.pop() for _ in range(n)]

    def _free_pages(self, pages):
        self._free.extend(reversed(pages))

    def new_kv(self, batch):
        return S.PagedKV(self, batch)

    def decode_state(self, batch):
        return self._ds.setdefault(batch, StubDecodeState(self, batch))

    def encode_images(self, pixels):
        return None

    def check_errors(self, sync=False):
        pass

    def text_forward(self, ids, feats, kv, logits="last"):
        B, N = ids.shape
        kv.reserve(N)
        kv.length = N
        out = torch.full((B, 1, 64), -1.0)
        for b in range(B):
            self.prefills.append(int(ids[b].sum()))
            out[b, 0, _stream(int(ids[b].sum()), 0)] = 1.0
        self.prefill_batches.append(B)
        return out


def _track_owners(cb, eng):
    """The stub model needs to know which request sits in which slot: record it when the scheduler seats a request."""
    orig_admit = cb._admit
    def admit(slots, reqs):
        before = dict(cb.running)
        orig_admit(slots, reqs)
        for slot, r in cb.running.items():
            if before.get(slot) is not r:
                eng.slot_owner[slot] = (int(r.input_ids.sum()), r.input_ids.shape[1])
    cb._admit = admit


def _patch_pick(monkeypatch):
    import pg_b200.generate as G
    monkeypatch.setattr(G, "_pick", lambda eng, logits, sample, step: logits.argmax(-1).to(torch.int64))


def test_scheduler_host_logic(monkeypatch):
    _patch_pick(monkeypatch)
    eng = StubEngine()
    cb = S.ContinuousBatcher(eng, slots=2, max_tokens=32, chunk=3)
    prompts = [torch.arange(1, 1 + n, dtype=torch.int64)[None] for n in (5, 9, 3, 6, 7)]
    budgets = [7, 2, 1, 9, 4]

    _track_owners(cb, eng)

    rids = [cb.submit(p, None, b) for p, b in zip(prompts, budgets)]
    done = cb.run()
    for rid, p, b in zip(rids, prompts, budgets):
        want = [_stream(int(p.sum()), t) for t in range(b)]
        assert done[rid].tokens == want, (rid, done[rid].tokens, want)
    assert not cb.running and not cb.queue and len(done) == 5
    assert all(n <= 3 for n in cb.ds.calls)                      # chunks never exceed the configured size
    assert eng.prefills == [int(p.sum()) for p in prompts]       # admitted in submission order, one prefill each
    parking = cb.kv.parking
    assert bool((cb.kv.page_table == parking).all())             # every row back on the parking page
    cb.close()
    assert sorted(eng._free) == list(range(64))                  # nothing leaked, nothing freed twice


def test_scheduler_eos_and_capacity(monkeypatch):
    _patch_pick(monkeypatch)
    eng = StubEngine()
    cb = S.ContinuousBatcher(eng, slots=1, max_tokens=16, chunk=4)
    _track_owners(cb, eng)
    p = torch.arange(1, 6, dtype=torch.int64)[None]
    stream = [_stream(int(p.sum()), t) for t in range(8)]
    eos = stream[4]
    rid = cb.submit(p, None, 8, eos_token_id=eos)
    done = cb.run()
    assert done[rid].tokens == stream[:stream.index(eos) + 1]
    cb.close()
    assert sorted(eng._free) == list(range(64))
    # a sequence that outgrows its slot is refused loudly
    cb = S.ContinuousBatcher(eng, slots=1, max_tokens=8, chunk=4)
    try:
        cb.submit(torch.arange(1, 7, dtype=torch.int64)[None], None, 10)     # 6 + 10 > 8: refused at submit()
        raise AssertionError("expected a capacity error")
    except ValueError as e:
        assert "capacity" in str(e)
    assert not cb.queue


def test_scheduler_history_ring_wraps(monkeypatch):
    """More decode steps than the history buffer holds (the stub's is 64 columns): the step counter keeps counting
    (it is also the sampler's RNG offset), the history is read as a ring, tokens stay right."""
    _patch_pick(monkeypatch)
    eng = StubEngine(pages=256)
    cb = S.ContinuousBatcher(eng, slots=2, max_tokens=256, chunk=5)
    _track_owners(cb, eng)
    prompts = [torch.arange(1, 1 + n, dtype=torch.int64)[None] for n in (4, 6, 5)]
    budgets = [150, 90, 70]
    rids = [cb.submit(p, None, b) for p, b in zip(prompts, budgets)]
    done = cb.run()
    for rid, p, b in zip(rids, prompts, budgets):
        assert done[rid].tokens == [_stream(int(p.sum()), t) for t in range(b)]
    assert int(cb.ds.step.item()) > cb.ds.max_hist                # wrapped at least once, never reset
    cb.close()
    assert sorted(eng._free) == list(range(256))



def test_scheduler_batches_prefills_of_equal_length(monkeypatch):
    """Queued requests with the same prompt length are admitted by ONE prefill call, as many as there are free slots;
    different lengths go in separate calls; every request still gets exactly its own token stream."""
    _patch_pick(monkeypatch)
    eng = StubEngine(pages=256)
    cb = S.ContinuousBatcher(eng, slots=4, max_tokens=64, chunk=4)
    _track_owners(cb, eng)
    lens = [6, 6, 6, 9, 6, 6, 9, 6]
    prompts = [(torch.arange(1, 1 + n, dtype=torch.int64) * (i + 1) % 97 + 1)[None] for i, n in enumerate(lens)]
    budgets = [5, 9, 3, 7, 4, 6, 2, 8]
    rids = [cb.submit(p, None, b) for p, b in zip(prompts, budgets)]
    done = cb.run()
    for rid, p, b in zip(rids, prompts, budgets):
        assert done[rid].tokens == [_stream(int(p.sum()), t) for t in range(b)], rid
    assert eng.prefill_batches[0] == 4                       # the four free slots took the first four length-6 prompts at once
    assert sum(eng.prefill_batches) == len(prompts) and len(eng.prefill_batches) < len(prompts)
    assert cb.prefill_calls == len(eng.prefill_batches)
    cb.close()
    assert sorted(eng._free) == list(range(256))
