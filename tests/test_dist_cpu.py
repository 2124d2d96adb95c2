"""CPU checks of the tensor-parallel host plumbing added in round 2 (no GPU, no torch.distributed):
the `pg_tp_exchange` ctypes mirror against the header, the exchange-buffer layout arithmetic, the lockstep driver that
emulates N ranks in one process, TP.run driving a launch generator, the vision batch shares and the static page table
of PagedKV."""
import ctypes as C
import os
import re

import pytest
import torch

from pg_b200 import dist as pgd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exchange_struct_mirrors_the_header():
    """Field order, types and size of pg_b200.dist.Exchange == `pg_tp_exchange` of include/pg_b200.h."""
    text = open(os.path.join(ROOT, "include", "pg_b200.h")).read()
    body = re.search(r"typedef struct pg_tp_exchange \{(.*?)\} pg_tp_exchange;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.rsplit(" ", 1)[0], decl
        base = "ptr" if "*" in decl else ("ll" if "long long" in decl else "int")
        for name in re.findall(r"\**\s*(\w+)\s*(?:,|$)", decl.split(" ", 1)[1] if base != "ll" else decl.split("long long", 1)[1]):
            if name not in ("void", "int", "const", "long"):
                fields.append((name, base))
    want = [(n, "ptr" if t is C.c_void_p else ("ll" if t is C.c_longlong else "int")) for n, t in pgd.Exchange._fields_]
    assert fields == want, (fields, want)
    assert C.sizeof(pgd.Exchange) == 4 * 8 + 2 * 8 + 4 * 4


def test_fabric_layout_arithmetic():
    size, D = 8, 2048
    total = pgd.Fabric.nbytes(size, D)
    x_slot, keys_slot = pgd.Fabric.MAX_ROWS * D * 8, pgd.Fabric.MAX_ROWS * 16
    assert total == 2 * size * (x_slot + keys_slot)
    buf = torch.zeros(total, dtype=torch.uint8)
    fab = pgd.Fabric(3, size, D, buf, peers_dev=1234)
    ex = fab.x(5, 38)
    assert (ex.region_off, ex.slot_bytes, ex.rank, ex.size, ex.index, ex.stride) == (0, x_slot, 3, size, 5, 38)
    k = fab.keys(37, 38)
    assert k.region_off == 2 * size * x_slot and k.slot_bytes == keys_slot          # keys region follows the x region
    assert k.region_off + 2 * size * keys_slot == total                            # ... and ends the buffer
    assert fab.x(5, 38) is ex                                                       # cached: the address handed to C stays valid
    assert not fab.lost_peer()
    fab._err_np[0] = 1
    assert fab.lost_peer()
    with pytest.raises(ValueError):
        pgd.Fabric(0, 9, D, buf, 0)


def test_shard_batch_covers_every_item_once():
    for n in (1, 7, 8, 9, 64, 65):
        for size in (1, 2, 4, 8):
            spans = [pgd.shard_batch(n, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            lens = [hi - lo for lo, hi in spans]
            assert max(lens) - min(lens) <= 1 and lens == sorted(lens, reverse=True)


def _rank_gen(rank, size, log):
    """A toy launch generator: one 'kernel', an all-reduce, another kernel, an all-gather; returns its rank."""
    x = torch.full((4,), float(rank + 1))
    log.append(("k0", rank))
    yield
    yield ("all_reduce", x)
    log.append(("k1", rank, x.clone()))
    yield
    out = torch.zeros(size, 4)
    yield ("all_gather", out, x * (rank + 1))
    log.append(("k2", rank, out.clone()))
    return rank


def test_lockstep_group_orders_ranks_and_performs_collectives():
    size, log = 3, []
    res = pgd.LockstepGroup().run([_rank_gen(r, size, log) for r in range(size)])
    assert res == [0, 1, 2]
    # kernel k of every rank is issued before kernel k+1 of any rank
    kinds = [e[0] for e in log]
    assert kinds == ["k0"] * size + ["k1"] * size + ["k2"] * size
    total = float(sum(r + 1 for r in range(size)))
    for e in log:
        if e[0] == "k1":
            assert torch.equal(e[2], torch.full((4,), total))            # sum in rank order, written back to every rank
        if e[0] == "k2":
            want = torch.stack([torch.full((4,), total * (r + 1)) for r in range(size)])
            assert torch.equal(e[2], want)                                  # concatenation in rank order


def test_lockstep_group_detects_divergence():
    def short():
        yield
    def long():
        yield
        yield
    with pytest.raises(RuntimeError, match="different launch sequences"):
        pgd.LockstepGroup().run([short(), long()])
    def reducer():
        yield ("all_reduce", torch.zeros(1))
    def plain():
        yield
    with pytest.raises(RuntimeError, match="diverged"):
        pgd.LockstepGroup().run([reducer(), plain()])


def test_tp_run_drives_a_generator_and_returns_its_value(monkeypatch):
    tp = pgd.TP(0, 2, None)
    calls = []
    monkeypatch.setattr(tp, "all_reduce", lambda t: calls.append(("ar", t)) or t)
    monkeypatch.setattr(tp, "all_gather", lambda out, t: calls.append(("ag", out, t)) or out)
    a, b = torch.zeros(1), torch.zeros(2, 1)

    def gen():
        yield
        yield ("all_reduce", a)
        yield ("all_gather", b, a)
        return 42
    assert tp.run(gen()) == 42 and [c[0] for c in calls] == ["ar", "ag"]
    assert pgd.TP().run((x for x in ())) is None                         # a generator that never yields
    with pytest.raises(RuntimeError, match="unknown collective"):
        tp.run(iter([("bogus",)]))
    emu = pgd.TP(0, 2, None, emulated=True)
    with pytest.raises(RuntimeError, match="LockstepGroup"):
        emu.all_reduce(torch.zeros(1))
    assert emu.make_fabric(64, "cpu") is None and pgd.TP().make_fabric(64, "cpu") is None


def test_paged_kv_table_keeps_its_address_while_the_sequence_grows():
    """The device page table of a sequence is allocated once and filled in place, so a decode graph captured over it
    survives cache growth (the e2e loop used to re-capture when the cache crossed a page-allocation boundary)."""
    from pg_b200.engine import PagedKV

    class Eng:
        page_size, device = 4, torch.device("cpu")

        def __init__(self):
            self._free = list(range(4095, -1, -1))

        def _alloc_pages(self, n):
            return [self._free.pop() for _ in range(n)]

        def _free_pages(self, pages):
            self._free.extend(reversed(pages))

    eng = Eng()
    kv = PagedKV(eng, batch=2)
    kv.reserve(10)
    ptr0, stride0 = kv.page_table.data_ptr(), kv.max_pages
    first = [list(p) for p in kv.pages]
    for total in (11, 40, 200, 1000):
        kv.reserve(total)
        assert kv.page_table.data_ptr() == ptr0 and kv.max_pages == stride0
        need = (total + 3) // 4
        assert all(len(p) >= need for p in kv.pages)
        assert [p[:len(f)] for p, f in zip(kv.pages, first)] == first             # pages never move
        for b in range(2):
            assert kv.page_table[b, :len(kv.pages[b])].tolist() == kv.pages[b]
    assert len({pg for p in kv.pages for pg in p}) == sum(len(p) for p in kv.pages)  # no page handed out twice
    kv.reserve(4 * (PagedKV.MIN_TABLE_PAGES + 5))                                   # beyond the initial table: it is regrown
    assert kv.max_pages > stride0 and kv.page_table[0, :len(kv.pages[0])].tolist() == kv.pages[0]
    kv.release()
    assert sorted(eng._free) == list(range(4096))
