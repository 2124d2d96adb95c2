"""Per-kernel parity through the C ABI against the CPU oracle / plain torch fp32 on the same
seeded inputs.  fp32 mode must match to 1e-4; bf16/fp16 are compared against the oracle run in
that dtype (same rounding points) with the north-star tolerance rtol 2e-2."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import paligemma_oracle as O  # noqa: E402
from pg_b200 import _cabi as cabi  # noqa: E402

DTYPES = [torch.float32, torch.bfloat16, torch.float16]
TOL = {torch.float32: dict(rtol=1e-4, atol=1e-4), torch.bfloat16: dict(rtol=2e-2, atol=2e-2),
       torch.float16: dict(rtol=4e-3, atol=4e-3)}


_KEEP = []


def dev(t):
    """Device copy kept alive until the test ends (raw pointers are handed to the C ABI, so a
    temporary must not be recycled by the caching allocator before the launch)."""
    d = t.cuda().contiguous()
    _KEEP.append(d)
    return d


@pytest.fixture(autouse=True)
def _release_device_copies():
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def gen(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(shape, generator=g) * scale).to(dtype)


def close(got, want, dtype, scale=1.0):
    tol = TOL[dtype]
    torch.testing.assert_close(got.float().cpu(), want.float(), rtol=tol["rtol"], atol=tol["atol"] * scale)


def st():
    return cabi.stream()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,D", [(5, 128), (3, 2048), (7, 1152), (300, 1152), (260, 2048), (65, 144)])
def test_rmsnorm_layernorm(dtype, rows, D):
    x, w, b = gen(rows, D, scale=3, dtype=dtype), gen(D, seed=1, scale=0.1, dtype=dtype), gen(D, seed=2, scale=0.1, dtype=dtype)
    out = torch.empty_like(dev(x))
    cabi.check(cabi.lib().pg_rmsnorm(out.data_ptr(), dev(x).data_ptr(), dev(w).data_ptr(), rows, D, 1e-6,
                                     cabi.DTYPE_CODE[dtype], st()))
    close(out, O.rms_norm(x, w, 1e-6), dtype)
    lw = (1 + w.float()).to(dtype)
    cabi.check(cabi.lib().pg_layernorm(out.data_ptr(), dev(x).data_ptr(), dev(lw).data_ptr(), dev(b).data_ptr(),
                                       rows, D, 1e-6, cabi.DTYPE_CODE[dtype], st()))
    close(out, F.layer_norm(x.float(), (D,), lw.float(), b.float(), 1e-6).to(dtype), dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_embed_merge(dtype):
    D, V, img_id, pad = 128, 300, 290, 0
    cfg = dict(hidden_size=D, image_token_index=img_id, pad_token_id=pad)
    emb = gen(V, D, scale=0.05, dtype=dtype)
    img = gen(2, 3, D, seed=3, dtype=dtype)
    ids = torch.tensor([[img_id, img_id, img_id, 2, 17, pad, 108], [img_id, img_id, img_id, 2, 299, 5, 108]])
    norm = float(torch.tensor(D ** 0.5, dtype=dtype).float())
    want = O.merge_embeddings(cfg, img, F.embedding(ids, emb), ids) * torch.tensor(D ** 0.5, dtype=dtype)
    out = torch.empty((ids.numel(), D), dtype=dtype, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    cabi.check(cabi.lib().pg_embed_merge(out.data_ptr(), dev(ids).data_ptr(), dev(emb).data_ptr(),
                                         dev(img.reshape(-1, D)).data_ptr(), ids.numel(), D, V, img_id, pad, 6,
                                         float(D ** 0.5), norm, err.data_ptr(), cabi.DTYPE_CODE[dtype], st()))
    assert int(err.item()) == 0
    torch.testing.assert_close(out.cpu().float(), want.reshape(-1, D).float(), rtol=1e-6, atol=1e-7)
    # an image token with no image row left flags the error the reference raises
    cabi.check(cabi.lib().pg_embed_merge(out.data_ptr(), dev(ids).data_ptr(), dev(emb).data_ptr(), None,
                                         ids.numel(), D, V, img_id, pad, 0, float(D ** 0.5), norm, err.data_ptr(),
                                         cabi.DTYPE_CODE[dtype], st()))
    assert int(err.item()) == 1


@pytest.mark.parametrize("dtype", DTYPES)
def test_im2col_patch_embed(dtype):
    B, C, S, p, Hv = 2, 3, 56, 14, 144
    px = gen(B, C, S, S, dtype=dtype)
    w, b = gen(Hv, C, p, p, seed=1, scale=0.05, dtype=dtype), gen(Hv, seed=2, scale=0.02, dtype=dtype)
    kc, kpad = C * p * p, 592
    col = torch.empty((B * 16, kpad), dtype=dtype, device="cuda")
    cabi.check(cabi.lib().pg_im2col(col.data_ptr(), dev(px).data_ptr(), B, C, S, S, p, kpad, cabi.DTYPE_CODE[dtype], st()))
    want_col = F.unfold(px.float(), p, stride=p).transpose(1, 2).reshape(B * 16, kc)
    assert torch.equal(col[:, :kc].float().cpu(), want_col)
    assert float(col[:, kc:].float().abs().max()) == 0.0
    wp = torch.zeros((Hv, kpad), dtype=dtype)
    wp[:, :kc] = w.reshape(Hv, kc)
    pos = gen(16, Hv, seed=5, scale=0.02, dtype=dtype)
    out = torch.empty((B * 16, Hv), dtype=dtype, device="cuda")
    cabi.check(cabi.lib().pg_gemm(out.data_ptr(), col.data_ptr(), dev(wp).data_ptr(), dev(b).data_ptr(),
                                  dev(pos).data_ptr(), B * 16, Hv, kpad, kpad, kpad, Hv, Hv, 16, cabi.EPI_BIAS_RES, 0, 1,
                                  cabi.DTYPE_CODE[dtype], st()))
    want = F.conv2d(px, w, b, stride=p).flatten(2).transpose(1, 2) + pos[None]
    close(out, want.reshape(B * 16, Hv), dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(20, 320, 128), (70, 272, 144), (33, 1152, 4304), (260, 2560, 2048)])
@pytest.mark.parametrize("epi", ["none", "bias", "bias_gelu", "bias_res", "res", "geglu", "f32"])
def test_gemm_simt_epilogues(dtype, M, N, K, epi):
    if K == 4304 and epi not in ("bias_res", "none"):
        pytest.skip("large-K case only for the fc2 shape")
    if K == 2048 and epi not in ("none", "geglu"):
        pytest.skip("prefill shape only for qkv / geglu")
    a = gen(M, K, dtype=dtype)
    s = 1.0 / math.sqrt(K)
    w, w2 = gen(N, K, seed=1, scale=s, dtype=dtype), gen(N, K, seed=2, scale=s, dtype=dtype)
    bias, res = gen(N, seed=3, scale=0.1, dtype=dtype), gen(M, N, seed=4, dtype=dtype)
    code = dict(none=cabi.EPI_NONE, bias=cabi.EPI_BIAS, bias_gelu=cabi.EPI_BIAS_GELU, bias_res=cabi.EPI_BIAS_RES,
                res=cabi.EPI_RES, geglu=cabi.EPI_GEGLU, f32=cabi.EPI_NONE)[epi]
    if epi == "none":
        want = F.linear(a, w)
    elif epi == "bias":
        want = F.linear(a, w, bias)
    elif epi == "bias_gelu":
        want = F.gelu(F.linear(a, w, bias), approximate="tanh")
    elif epi == "bias_res":
        want = F.linear(a, w, bias) + res
    elif epi == "res":
        want = F.linear(a, w) + res
    elif epi == "geglu":
        want = F.gelu(F.linear(a, w), approximate="tanh") * F.linear(a, w2)
    else:
        want = F.linear(a, w).float()
    wd = dev(torch.cat([w, w2], 0)) if epi == "geglu" else dev(w)
    out = torch.empty((M, N), dtype=torch.float32 if epi == "f32" else dtype, device="cuda")
    cabi.check(cabi.lib().pg_gemm(out.data_ptr(), dev(a).data_ptr(), wd.data_ptr(), dev(bias).data_ptr(),
                                  dev(res).data_ptr(), M, N, K, K, K, N, N, 0, code, 1 if epi == "f32" else 0, 1,
                                  cabi.DTYPE_CODE[dtype], st()))
    close(out, want, dtype)


def _paged(B, T, nkv, hd, dtype, page=16, seed=0):
    """Random K/V (B,nkv,T,hd) scattered into a shuffled page pool."""
    k, v = gen(B, nkv, T, hd, seed=seed, dtype=dtype), gen(B, nkv, T, hd, seed=seed + 1, dtype=dtype)
    npg = (T + page - 1) // page + 1
    perm = torch.randperm(B * npg + 3, generator=torch.Generator().manual_seed(5))[: B * npg].view(B, npg)
    kp = torch.zeros((B * npg + 3, page, nkv * hd), dtype=dtype)
    vp = torch.zeros_like(kp)
    for b in range(B):
        for t in range(T):
            kp[perm[b, t // page], t % page] = k[b, :, t].reshape(-1)
            vp[perm[b, t // page], t % page] = v[b, :, t].reshape(-1)
    return k, v, dev(kp), dev(vp), dev(perm.to(torch.int32)), npg, page


def _ref_attention(q, k, v, scale_div, dtype):
    """(B,H,q,hd) x (B,Hkv,T,hd): the oracle's rounding points (gemma_attention)."""
    rep = q.shape[1] // k.shape[1]
    k = k.repeat_interleave(rep, 1)
    v = v.repeat_interleave(rep, 1)
    w = torch.matmul(q, k.transpose(2, 3)) / scale_div
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(dtype)
    return torch.matmul(w, v)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,T,nq,hd", [(1, 5, 4, 32), (2, 77, 8, 256), (3, 300, 8, 256), (1, 1344, 8, 256)])
def test_decode_attention_paged(dtype, B, T, nq, hd):
    nkv = 1
    k, v, kp, vp, pt, npg, page = _paged(B, T, nkv, hd, dtype)
    q = gen(B, nq, 1, hd, seed=9, dtype=dtype)
    want = _ref_attention(q, k, v, math.sqrt(hd), dtype).transpose(1, 2).reshape(B, nq * hd)
    out = torch.empty((B, nq * hd), dtype=dtype, device="cuda")
    splits = 16
    ws = torch.zeros(int(cabi.lib().pg_decode_attention_ws_floats(B, nq, hd, splits)), device="cuda")
    cnt = torch.zeros(B * nkv, dtype=torch.int32, device="cuda")
    kvl = torch.full((B,), T - 1, dtype=torch.int32, device="cuda")
    qd = dev(q.transpose(1, 2).reshape(B, nq * hd))
    for _ in range(2):  # second launch checks that the counters were re-armed
        cabi.check(cabi.lib().pg_decode_attention(out.data_ptr(), qd.data_ptr(), kp.data_ptr(), vp.data_ptr(),
                                                  pt.data_ptr(), npg, page, kvl.data_ptr(), 1, B, nq, nkv, hd,
                                                  float(math.sqrt(hd)), ws.data_ptr(), cnt.data_ptr(), splits,
                                                  cabi.DTYPE_CODE[dtype], st()))
    close(out, want, dtype)
    assert int(cnt.abs().sum().item()) == 0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,q,T,nq,nkv,hd,paged", [(2, 16, 16, 2, 2, 72, False), (1, 20, 20, 4, 1, 32, True),
                                                   (2, 37, 70, 8, 1, 256, True), (1, 256, 256, 16, 16, 72, False)])
def test_attention_general(dtype, B, q, T, nq, nkv, hd, paged):
    qq = gen(B, nq, q, hd, seed=9, dtype=dtype)
    out = torch.empty((B * q, nq * hd), dtype=dtype, device="cuda")
    qd = dev(qq.transpose(1, 2).reshape(B * q, nq * hd))
    if paged:
        k, v, kp, vp, pt, npg, page = _paged(B, T, nkv, hd, dtype)
        kvl = torch.full((B,), T - q, dtype=torch.int32, device="cuda")
        want = _ref_attention(qq, k, v, math.sqrt(hd), dtype)
        cabi.check(cabi.lib().pg_attention(out.data_ptr(), nq * hd, qd.data_ptr(), nq * hd, kp.data_ptr(), vp.data_ptr(),
                                           0, 0, pt.data_ptr(), npg, page, kvl.data_ptr(), 0, q, B, q, nq, nkv, hd,
                                           float(math.sqrt(hd)), 1, cabi.DTYPE_CODE[dtype], st()))
    else:
        k, v = gen(B, nkv, T, hd, seed=1, dtype=dtype), gen(B, nkv, T, hd, seed=2, dtype=dtype)
        w = torch.matmul(qq, k.transpose(2, 3)) * (hd ** -0.5)
        want = torch.matmul(F.softmax(w, dim=-1, dtype=torch.float32).to(dtype), v)
        # SigLIP layout: one fused [tokens, 3*H] buffer
        H = nq * hd
        fused = torch.zeros((B * T, 3 * H), dtype=dtype)
        fused[:, :H] = qq.transpose(1, 2).reshape(B * q, H)
        fused[:, H:2 * H] = k.transpose(1, 2).reshape(B * T, H)
        fused[:, 2 * H:] = v.transpose(1, 2).reshape(B * T, H)
        fd = dev(fused)
        cabi.check(cabi.lib().pg_attention(out.data_ptr(), H, fd.data_ptr(), 3 * H, fd[:, H:].data_ptr(),
                                           fd[:, 2 * H:].data_ptr(), 3 * H, T * 3 * H, None, 0, 0, None, T, 0, B, q, nq,
                                           nkv, hd, float(hd ** -0.5), 0, cabi.DTYPE_CODE[dtype], st()))
    close(out, want.transpose(1, 2).reshape(B * q, nq * hd), dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B", [1, 3, 8])
@pytest.mark.parametrize("D,nq,hd,F_", [(128, 4, 32, 512), (2048, 8, 256, 16384)])
def test_decode_layer_kernels(dtype, B, D, nq, hd, F_):
    """decode_qkv (+RoPE +append), gemv_res, decode_gateup against the oracle's layer pieces."""
    if D == 2048 and B == 3:
        pytest.skip("covered by B=1 and B=8")
    nkv, page, T0 = 1, 16, 21
    t = dict(num_attention_heads=nq, num_key_value_heads=nkv, head_dim=hd)
    x = gen(B, D, scale=2, dtype=dtype)
    s = 1.0 / math.sqrt(D)
    lnw = gen(D, seed=1, scale=0.1, dtype=dtype)
    wq, wk, wv = gen(nq * hd, D, seed=2, scale=s, dtype=dtype), gen(hd, D, seed=3, scale=s, dtype=dtype), gen(hd, D, seed=4, scale=s, dtype=dtype)
    pos = torch.tensor([T0 + 1 + 3 * b for b in range(B)], dtype=torch.int32)
    h = O.rms_norm(x, lnw, 1e-6)
    q = F.linear(h, wq).view(B, 1, nq, hd).transpose(1, 2)
    k = F.linear(h, wk).view(B, 1, nkv, hd).transpose(1, 2)
    v = F.linear(h, wv)
    cos, sin = O.rope_cos_sin(pos[:, None], hd, 10000.0, 8192, dtype)
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    q_ref = (q * cos) + (O._rot_half(q) * sin)
    k_ref = (k * cos) + (O._rot_half(k) * sin)
    npg = 4
    kp = torch.zeros((B * npg, page, nkv * hd), dtype=dtype, device="cuda")
    vp = torch.zeros_like(kp)
    pt = dev(torch.arange(B * npg, dtype=torch.int32).view(B, npg))
    kvl = torch.full((B,), T0, dtype=torch.int32, device="cuda")
    inv = dev(O.inv_freq(hd, 10000.0, dtype))
    qo = torch.empty((B, nq * hd), dtype=dtype, device="cuda")
    wqkv = dev(torch.cat([wq, wk, wv], 0))
    cabi.check(cabi.lib().pg_decode_qkv(qo.data_ptr(), dev(x).data_ptr(), dev(lnw).data_ptr(), wqkv.data_ptr(),
                                        inv.data_ptr(), dev(pos).data_ptr(), kp.data_ptr(), vp.data_ptr(), pt.data_ptr(),
                                        npg, page, kvl.data_ptr(), B, D, nq, nkv, hd, 1e-6, 8192, None, None,
                                        cabi.DTYPE_CODE[dtype], st()))
    close(qo, q_ref.transpose(1, 2).reshape(B, nq * hd), dtype)
    got_k = torch.stack([kp[pt[b, T0 // page].item(), T0 % page] for b in range(B)])
    got_v = torch.stack([vp[pt[b, T0 // page].item(), T0 % page] for b in range(B)])
    close(got_k, k_ref.reshape(B, hd), dtype)
    close(got_v, v, dtype)
    # o_proj-shaped and down_proj-shaped GEMV with residual
    for K in (nq * hd, F_):
        a = gen(B, K, seed=7, dtype=dtype)
        w = gen(D, K, seed=8, scale=1 / math.sqrt(K), dtype=dtype)
        out = torch.empty((B, D), dtype=dtype, device="cuda")
        cabi.check(cabi.lib().pg_gemv_res(out.data_ptr(), dev(a).data_ptr(), dev(w).data_ptr(), dev(x).data_ptr(), B, D, K,
                                          None, cabi.DTYPE_CODE[dtype], st()))
        close(out, x + F.linear(a, w), dtype)
    # gate/up + GeGLU with the post-attention norm fused
    wg, wu = gen(F_, D, seed=10, scale=s, dtype=dtype), gen(F_, D, seed=11, scale=s, dtype=dtype)
    out = torch.empty((B, F_), dtype=dtype, device="cuda")
    cabi.check(cabi.lib().pg_decode_gateup(out.data_ptr(), dev(x).data_ptr(), dev(lnw).data_ptr(),
                                           dev(torch.cat([wg, wu], 0)).data_ptr(), B, D, F_, 1e-6, None, None,
                                           cabi.DTYPE_CODE[dtype], st()))
    close(out, F.gelu(F.linear(h, wg), approximate="tanh") * F.linear(h, wu), dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,D,V", [(1, 128, 1280), (4, 2048, 8065)])
def test_lmhead_argmax_and_step_advance(dtype, B, D, V):
    x, lnw = gen(B, D, scale=2, dtype=dtype), gen(D, seed=1, scale=0.1, dtype=dtype)
    emb = gen(V, D, seed=2, scale=0.05, dtype=dtype)
    want = F.linear(O.rms_norm(x, lnw, 1e-6), emb).float()
    logits = torch.empty((B, V), dtype=torch.float32, device="cuda")
    keys = torch.zeros(B, dtype=torch.int64, device="cuda")
    cabi.check(cabi.lib().pg_decode_lmhead(logits.data_ptr(), dev(x).data_ptr(), dev(lnw).data_ptr(), dev(emb).data_ptr(),
                                           B, D, V, 1e-6, keys.data_ptr(), None, None, cabi.DTYPE_CODE[dtype], st()))
    close(logits, want, dtype)
    ids = torch.zeros(B, dtype=torch.int64, device="cuda")
    hist = torch.zeros((B, 4), dtype=torch.int64, device="cuda")
    step = torch.full((1,), 2, dtype=torch.int32, device="cuda")
    kvl = torch.full((B,), 9, dtype=torch.int32, device="cuda")
    pos = torch.full((B,), 11, dtype=torch.int32, device="cuda")
    cabi.check(cabi.lib().pg_step_advance(ids.data_ptr(), hist.data_ptr(), 4, step.data_ptr(), keys.data_ptr(), None,
                                          kvl.data_ptr(), pos.data_ptr(), B, None, st()))
    # the kernel's own logits decide (ties -> lowest index, as torch.argmax on the same values)
    assert ids.cpu().tolist() == torch.argmax(logits.cpu(), dim=-1).tolist()
    assert hist[:, 2].cpu().tolist() == ids.cpu().tolist() and int(step.item()) == 3
    assert kvl.cpu().tolist() == [10] * B and pos.cpu().tolist() == [12] * B and int(keys.abs().sum().item()) == 0
    # standalone argmax with planted ties
    lg = logits.clone()
    lg[:, 7] = lg.max() + 1
    lg[:, 3] = lg[:, 7]
    out = torch.empty(B, dtype=torch.int64, device="cuda")
    cabi.check(cabi.lib().pg_argmax(out.data_ptr(), lg.data_ptr(), keys.data_ptr(), B, V, st()))
    assert out.cpu().tolist() == [3] * B


@pytest.mark.parametrize("V", [1280, 257216])
def test_top_p_nucleus_matches_reference_rule(V):
    B, temp, top_p = 3, 0.8, 0.9
    logits = gen(B, V, scale=2.3)
    logits[1] = logits[1].to(torch.bfloat16).float()  # many exact ties, as bf16 logits have
    logits[2, :5] += 12.0                                # a peaked row: nucleus of a few tokens
    dist = O.top_p_distribution(logits, temp, top_p)
    want_count = (dist > 0).sum(-1)
    out = torch.empty(B, dtype=torch.int64, device="cuda")
    probs = torch.empty((B, V), dtype=torch.float32, device="cuda")
    nuc = torch.zeros(B, dtype=torch.int32, device="cuda")
    off = torch.zeros(1, dtype=torch.int32, device="cuda")
    ld = dev(logits)
    draws = []
    for i in range(200):
        off.fill_(i)
        cabi.check(cabi.lib().pg_top_p_sample(out.data_ptr(), ld.data_ptr(), probs.data_ptr(), B, V, temp, top_p, 1234,
                                              off.data_ptr(), nuc.data_ptr(), st()))
        draws.append(out.cpu().clone())
    torch.testing.assert_close(probs.cpu(), torch.softmax(logits / temp, -1), rtol=1e-4, atol=1e-9)
    # nucleus size: identical up to fp32 summation-order noise at the boundary
    for b in range(B):
        assert abs(int(nuc[b]) - int(want_count[b])) <= max(2, int(0.001 * int(want_count[b]))), (nuc, want_count)
    draws = torch.stack(draws)  # (200, B)
    # every draw lies in (a hair's breadth of) the reference nucleus
    thresh = torch.where(dist > 0, dist, torch.ones_like(dist)).min(-1).values
    p_ref = torch.softmax(logits / temp, -1)
    for b in range(B):
        pd = p_ref[b][draws[:, b]]
        assert bool((pd >= p_ref[b][dist[b] > 0].min() * (1 - 1e-4)).all())
    # the peaked row is dominated by its few boosted tokens
    assert set(draws[:, 2].tolist()) <= set(torch.nonzero(dist[2]).flatten().tolist())
    assert len(set(draws[:, 0].tolist())) > 10  # and it really samples


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (260, 2560, 2048), (256, 1152, 4304), (70, 272, 144),
                                   (33, 4304, 1152), (512, 1152, 592), (1000, 2048, 2048), (16, 8, 8)])
@pytest.mark.parametrize("epi", ["none", "bias", "bias_gelu", "bias_res", "res", "geglu", "f32", "posemb"])
def test_gemm_tcgen05(dtype, M, N, K, epi):
    """The tcgen05/TMEM/TMA GEMM (impl=2) against torch on CPU in the same dtype and against the SIMT
    kernel (same rounding points, different accumulation order)."""
    a = gen(M, K, dtype=dtype)
    s = 1.0 / math.sqrt(K)
    w, w2 = gen(N, K, seed=1, scale=s, dtype=dtype), gen(N, K, seed=2, scale=s, dtype=dtype)
    bias = gen(N, seed=3, scale=0.1, dtype=dtype)
    res_mod = 16 if epi == "posemb" else 0
    res = gen(16 if epi == "posemb" else M, N, seed=4, dtype=dtype)
    code = dict(none=cabi.EPI_NONE, bias=cabi.EPI_BIAS, bias_gelu=cabi.EPI_BIAS_GELU, bias_res=cabi.EPI_BIAS_RES,
                res=cabi.EPI_RES, geglu=cabi.EPI_GEGLU, f32=cabi.EPI_NONE, posemb=cabi.EPI_BIAS_RES)[epi]
    if epi == "none":
        want = F.linear(a, w)
    elif epi == "bias":
        want = F.linear(a, w, bias)
    elif epi == "bias_gelu":
        want = F.gelu(F.linear(a, w, bias), approximate="tanh")
    elif epi == "bias_res":
        want = F.linear(a, w, bias) + res
    elif epi == "res":
        want = F.linear(a, w) + res
    elif epi == "geglu":
        want = F.gelu(F.linear(a, w), approximate="tanh") * F.linear(a, w2)
    elif epi == "posemb":
        want = F.linear(a, w, bias) + res[torch.arange(M) % 16]
    else:
        want = F.linear(a, w).float()
    wd = dev(torch.cat([w, w2], 0)) if epi == "geglu" else dev(w)
    ad, bd, rd = dev(a), dev(bias), dev(res)
    outs = []
    for impl in (2, 1):
        out = torch.full((M, N), float("nan"), dtype=torch.float32 if epi == "f32" else dtype, device="cuda")
        cabi.check(cabi.lib().pg_gemm(out.data_ptr(), ad.data_ptr(), wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), M, N, K,
                                      K, K, N, N, res_mod, code, 1 if epi == "f32" else 0, impl,
                                      cabi.DTYPE_CODE[dtype], st()))
        outs.append(out)
    close(outs[0], want, dtype)
    close(outs[0], outs[1].cpu(), dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,S,heads,hd,hs", [(2, 256, 16, 72, 128), (1, 16, 2, 72, 128), (3, 64, 4, 72, 128), (1, 300, 2, 128, 128),
                                              (2, 256, 16, 72, 80), (1, 16, 2, 72, 80), (3, 64, 4, 72, 80), (2, 200, 3, 72, 80),
                                              (1, 130, 2, 64, 64), (2, 256, 2, 32, 32), (1, 300, 2, 72, 128), (2, 256, 4, 80, 80)])
def test_attention_tc_siglip_padded_layout(dtype, B, S, heads, hd, hs):
    """tcgen05 attention on the SigLIP layout: fused [tokens, 3*heads*hs] with each head padded to hs columns
    (<= 256 keys and head_dim <= 80 take the ViT kernel, the rest the tiled kernel)."""
    q, k, v = (gen(B, heads, S, hd, seed=i, dtype=dtype) for i in (1, 2, 3))
    w = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
    want = torch.matmul(F.softmax(w, dim=-1, dtype=torch.float32).to(dtype), v).transpose(1, 2).reshape(B * S, heads * hd)
    fused = torch.zeros((B * S, 3, heads, hs), dtype=dtype)
    for i, t in enumerate((q, k, v)):
        fused[:, i, :, :hd] = t.transpose(1, 2).reshape(B * S, heads, hd)
    fd = dev(fused.reshape(B * S, 3 * heads * hs))
    out = torch.full((B * S, heads * hd), float("nan"), dtype=dtype, device="cuda")
    ld = 3 * heads * hs
    cabi.check(cabi.lib().pg_attention_tc(out.data_ptr(), heads * hd, fd.data_ptr(), B * S, ld, 0, fd.data_ptr(), fd.data_ptr(),
                                          B * S, ld, heads * hs, 2 * heads * hs, hs, S, None, 0, 0, None, S, 0, B, S, heads,
                                          heads, hd, float(hd ** -0.5), 0, cabi.DTYPE_CODE[dtype], st()))
    close(out, want, dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,q,T,nq", [(1, 260, 260, 8), (2, 37, 70, 8), (1, 130, 516, 4), (2, 16, 16, 2)])
def test_attention_tc_gemma_paged(dtype, B, q, T, nq):
    """tcgen05 attention over the paged KV pool (page == 64-key tile), MQA, head_dim 256."""
    hd, nkv = 256, 1
    qq = gen(B, nq, q, hd, seed=9, dtype=dtype)
    k, v, kp, vp, pt, npg, page = _paged(B, T, nkv, hd, dtype, page=64)
    want = _ref_attention(qq, k, v, math.sqrt(hd), dtype).transpose(1, 2).reshape(B * q, nq * hd)
    qd = dev(qq.transpose(1, 2).reshape(B * q, nq * hd))
    kvl = torch.full((B,), T - q, dtype=torch.int32, device="cuda")
    out = torch.full((B * q, nq * hd), float("nan"), dtype=dtype, device="cuda")
    rows = kp.shape[0] * page
    cabi.check(cabi.lib().pg_attention_tc(out.data_ptr(), nq * hd, qd.data_ptr(), B * q, nq * hd, 0, kp.data_ptr(), vp.data_ptr(),
                                          rows, nkv * hd, 0, 0, hd, 0, pt.data_ptr(), npg, page, kvl.data_ptr(), 0, q, B, q, nq,
                                          nkv, hd, float(math.sqrt(hd)), 1, cabi.DTYPE_CODE[dtype], st()))
    close(out, want, dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(4096, 4304, 1152), (4100, 2560, 512), (16384, 1152, 4304)])
@pytest.mark.parametrize("epi", ["bias", "bias_gelu", "bias_res", "f32"])
def test_gemm_tcgen05_large_m_multicast_pairs(dtype, M, N, K, epi):
    """Large-M shapes take the 128x256 tiles with 2-CTA clusters and TMA-multicast W halves (odd m-tile counts
    exercise the ghost tile of the last pair)."""
    a = gen(M, K, dtype=dtype)
    w = gen(N, K, seed=1, scale=1.0 / math.sqrt(K), dtype=dtype)
    bias, res = gen(N, seed=3, scale=0.1, dtype=dtype), gen(M, N, seed=4, dtype=dtype)
    code = dict(bias=cabi.EPI_BIAS, bias_gelu=cabi.EPI_BIAS_GELU, bias_res=cabi.EPI_BIAS_RES, f32=cabi.EPI_NONE)[epi]
    ad, wd, bd, rd = dev(a), dev(w), dev(bias), dev(res)
    if epi == "bias":
        want = F.linear(a, w, bias)
    elif epi == "bias_gelu":
        want = F.gelu(F.linear(a, w, bias), approximate="tanh")
    elif epi == "bias_res":
        want = F.linear(a, w, bias) + res
    else:
        want = F.linear(a, w).float()
    out = torch.full((M, N), float("nan"), dtype=torch.float32 if epi == "f32" else dtype, device="cuda")
    cabi.check(cabi.lib().pg_gemm(out.data_ptr(), ad.data_ptr(), wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), M, N, K, K, K,
                                  N, N, 0, code, 1 if epi == "f32" else 0, 2, cabi.DTYPE_CODE[dtype], st()))
    torch.cuda.synchronize()
    close(out, want, dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M", [4, 8, 16, 32, 33, 64, 100])
@pytest.mark.parametrize("N,K,epi", [(2560, 2048, "none"), (2048, 2048, "res"), (2048, 16384, "res"), (16384, 2048, "geglu"),
                                     (8064, 2048, "f32"), (4304, 1152, "bias_gelu"), (1152, 4304, "bias_res")])
def test_gemm_tcgen05_skinny_swap_ab(dtype, M, N, K, epi, monkeypatch):
    """Batched-decode shapes (16..128 token rows): the swap-AB weight-streaming kernel, with and without the
    cluster split along K, against torch in the same dtype."""
    if dtype == torch.float16 and M not in (8, 32, 33):
        pytest.skip("fp16 on three row counts only")
    a = gen(M, K, dtype=dtype)
    s = 1.0 / math.sqrt(K)
    w, w2 = gen(N, K, seed=1, scale=s, dtype=dtype), gen(N, K, seed=2, scale=s, dtype=dtype)
    bias, res = gen(N, seed=3, scale=0.1, dtype=dtype), gen(M, N, seed=4, dtype=dtype)
    code = dict(none=cabi.EPI_NONE, res=cabi.EPI_RES, geglu=cabi.EPI_GEGLU, f32=cabi.EPI_NONE, bias_gelu=cabi.EPI_BIAS_GELU,
                bias_res=cabi.EPI_BIAS_RES)[epi]
    if epi == "none":
        want = F.linear(a, w)
    elif epi == "res":
        want = F.linear(a, w) + res
    elif epi == "geglu":
        want = F.gelu(F.linear(a, w), approximate="tanh") * F.linear(a, w2)
    elif epi == "bias_gelu":
        want = F.gelu(F.linear(a, w, bias), approximate="tanh")
    elif epi == "bias_res":
        want = F.linear(a, w, bias) + res
    else:
        want = F.linear(a, w).float()
    wd = dev(torch.cat([w, w2], 0)) if epi == "geglu" else dev(w)
    ad, bd, rd = dev(a), dev(bias), dev(res)
    out = torch.full((M, N), float("nan"), dtype=torch.float32 if epi == "f32" else dtype, device="cuda")
    cabi.check(cabi.lib().pg_gemm(out.data_ptr(), ad.data_ptr(), wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), M, N, K, K, K,
                                  N, N, 0, code, 1 if epi == "f32" else 0, 2, cabi.DTYPE_CODE[dtype], st()))
    torch.cuda.synchronize()
    close(out, want, dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M", [8, 32, 100, 129, 144, 260, 272, 300, 512])
@pytest.mark.parametrize("N,K,epi", [(2560, 2048, "none"), (2048, 2048, "res"), (2048, 16384, "res"), (16384, 2048, "geglu"),
                                     (768, 2048, "none"), (1000, 1088, "res"), (2048, 4096, "geglu")])
def test_gemm_tcgen05_prompt_rows_swapped_pairs(dtype, M, N, K, epi):
    """Prompt-sized row counts (129..512: the 260-token prefill, the cache-off recompute up to 512 tokens; also the
    batched-decode rows 4..128 for the residual projections) take the
    CTA-pair kernel with the WEIGHTS as the M = 256 operand and all tokens in one or two accumulators, split along K
    over the pairs, plus the reduce / epilogue pass (csrc/gemm_tcgen05_swap.cu).  Shapes: the Gemma projections at
    tp 1 / tp 8 (768-row q/k/v shard) and ragged N / K tails.  Against torch in the same dtype."""
    if dtype == torch.float16 and M not in (260, 512):
        pytest.skip("fp16 on two row counts only")
    ws = torch.empty(80 << 20, dtype=torch.uint8, device="cuda")
    cabi.check(cabi.lib().pg_set_workspace(ws.data_ptr(), ws.numel()))
    a = gen(M, K, dtype=dtype)
    s = 1.0 / math.sqrt(K)
    w, w2 = gen(N, K, seed=1, scale=s, dtype=dtype), gen(N, K, seed=2, scale=s, dtype=dtype)
    res = gen(M, N, seed=4, dtype=dtype)
    code = dict(none=cabi.EPI_NONE, res=cabi.EPI_RES, geglu=cabi.EPI_GEGLU)[epi]
    if epi == "none":
        want = F.linear(a, w)
    elif epi == "res":
        want = F.linear(a, w) + res
    else:
        want = F.gelu(F.linear(a, w), approximate="tanh") * F.linear(a, w2)
    wd = dev(torch.cat([w, w2], 0)) if epi == "geglu" else dev(w)
    ad, rd = dev(a), dev(res)
    out = torch.full((M, N), float("nan"), dtype=dtype, device="cuda")
    before = cabi.launch_count()
    cabi.check(cabi.lib().pg_gemm(out.data_ptr(), ad.data_ptr(), wd.data_ptr(), None, rd.data_ptr(), M, N, K, K, K, N, N, 0,
                                  code, 0, 3, cabi.DTYPE_CODE[dtype], st()))
    torch.cuda.synchronize()
    assert cabi.launch_count() - before == 2                       # the pair GEMM + its reduce / epilogue pass
    close(out, want, dtype)
    # the same problem through the row-major tcgen05 kernels (PG_GEMM_SWAP off is an env switch read once, so compare
    # with the SIMT path instead: same rounding points, different accumulation order)
    ref = torch.empty_like(out)
    cabi.check(cabi.lib().pg_gemm(ref.data_ptr(), ad.data_ptr(), wd.data_ptr(), None, rd.data_ptr(), M, N, K, K, K, N, N, 0,
                                  code, 0, 1, cabi.DTYPE_CODE[dtype], st()))
    close(out, ref.cpu(), dtype)
