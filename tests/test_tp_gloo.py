"""World-size-2 (gloo, CPU) checks of the tensor-parallel host logic: the shard layout, the
residual-once rule and the vocab-shard argmax combine reproduce the unsharded layer."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from pg_b200 import dist as pgd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tp = pgd.TP(rank, world, None)
        g = torch.Generator().manual_seed(0)
        D, nq, hd, Fi, V, B = 64, 4, 16, 96, 40, 3
        q, k, v = torch.randn(nq * hd, D, generator=g), torch.randn(hd, D, generator=g), torch.randn(hd, D, generator=g)
        o, gate, up, down = (torch.randn(D, nq * hd, generator=g), torch.randn(Fi, D, generator=g),
                             torch.randn(Fi, D, generator=g), torch.randn(D, Fi, generator=g))
        emb = torch.randn(V, D, generator=g)
        x = torch.randn(B, D, generator=g)
        att_full = torch.randn(B, nq * hd, generator=g)        # stand-in for the per-head attention output
        qkv_l, o_l, gu_l, down_l = pgd.shard_text_layer(q, k, v, o, gate, up, down, rank, world)
        nql = nq // world
        # the q rows of this rank are its heads; k, v replicated
        assert torch.equal(qkv_l[: nql * hd], q[rank * nql * hd:(rank + 1) * nql * hd])
        assert torch.equal(qkv_l[nql * hd: nql * hd + hd], k) and torch.equal(qkv_l[nql * hd + hd:], v)
        # o_proj: partial sums over this rank's heads, residual on rank 0 only, then all-reduce
        att_l = att_full[:, rank * nql * hd:(rank + 1) * nql * hd]
        part = F.linear(att_l, o_l) + (x if rank == 0 else 0)
        tp.all_reduce(part)
        want = x + F.linear(att_full, o)
        assert torch.allclose(part, want, rtol=1e-5, atol=1e-5)
        # MLP: column-split gate/up, row-split down
        Fl = Fi // world
        gl = F.gelu(F.linear(part, gu_l[:Fl]), approximate="tanh") * F.linear(part, gu_l[Fl:])
        y = F.linear(gl, down_l) + (part if rank == 0 else 0)
        tp.all_reduce(y)
        want2 = want + F.linear(F.gelu(F.linear(want, gate), approximate="tanh") * F.linear(want, up), down)
        assert torch.allclose(y, want2, rtol=1e-4, atol=1e-4)
        # vocab-sharded lm_head with ties across shards: lowest global index wins, like torch.argmax
        Vl = V // world
        emb_t = emb.clone()
        emb_t[Vl + 3] = emb_t[2]                                 # identical rows in different shards -> tied logits
        logits_l = F.linear(y, pgd.shard_rows(emb_t, rank, world))
        def key(val, idx):
            b = val.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
            b = torch.where(b & 0x80000000 != 0, (~b) & 0xFFFFFFFF, b | 0x80000000)
            packed = (b << 32) | (0xFFFFFFFF - idx)
            return packed
        best_val, best_idx = logits_l.max(-1)
        # first maximal local index, as the kernel's packed atomicMax yields
        best_idx = (logits_l == best_val[:, None]).float().argmax(-1)
        keys = key(best_val.contiguous(), best_idx)
        gathered = torch.zeros(world, B, dtype=torch.int64)
        tp.all_gather(gathered, keys)
        tok = pgd.combine_argmax_keys(gathered, Vl)
        full = torch.zeros(world, B, Vl)
        tp.all_gather(full, logits_l)
        want_tok = full.permute(1, 0, 2).reshape(B, V).argmax(-1)
        assert tok.tolist() == want_tok.tolist(), (tok, want_tok)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_tensor_parallel_host_logic_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_combine_argmax_keys_negative_values():
    # all logits negative: ordered-bits mapping must still pick the largest
    vals = torch.tensor([[-3.0, -1.0], [-2.0, -1.0]])  # [size=2, B=2]
    idx = torch.tensor([[5, 7], [1, 0]])
    b = vals.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    b = torch.where(b & 0x80000000 != 0, (~b) & 0xFFFFFFFF, b | 0x80000000)
    keys = (b << 32) | (0xFFFFFFFF - idx)
    tok = pgd.combine_argmax_keys(keys, 100)
    assert tok.tolist() == [101, 7]   # row 0: -2 on rank 1 (global 101); row 1: tie -1 -> lowest global index 7
