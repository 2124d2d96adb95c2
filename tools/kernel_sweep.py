#!/usr/bin/env python
"""Per-kernel timing of the bs-1 decode step on the full-size model (random bf16 weights made on
the GPU: timing only).  Every kernel is launched over all 18 layers' weights in turn so no launch
finds its weights in L2; reports median us/launch and GB/s, then the whole CUDA-graph step.
Tunables are read from the environment by libpg_b200 (PG_PDL, PG_GEMV_CTAS_PER_SM, PG_DOWN_KS)."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200"))
import torch  # noqa: E402
from pg_b200 import synth, _cabi as cabi  # noqa: E402
from pg_b200.engine import PaliGemmaEngine  # noqa: E402


def gpu_weights(cfg, dtype):
    g = torch.Generator(device="cuda").manual_seed(1)
    sd = {}
    for key, shape, kind in synth.state_dict_spec(cfg):
        t = torch.randn(shape, generator=g, device="cuda", dtype=torch.float32) * synth._STD[kind]
        if kind == "ln_w":
            t += 1
        sd[key] = t.to(dtype)
    return sd


def timeit(fn, n_inner, reps=7):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(1e3 * e0.elapsed_time(e1) / n_inner)
    return statistics.median(out), min(out)


def main():
    B = int(os.environ.get("SWEEP_B", "1"))
    T = int(os.environ.get("SWEEP_T", "300"))
    cfg = synth.CONFIGS["paligemma-3b-pt-224"]
    tp_size = int(os.environ.get("SWEEP_TP", "1"))      # > 1: rank 0's SHARD shapes, kernels timed without the exchange
    tp = None
    if tp_size > 1:
        from pg_b200.dist import TP, Fabric
        tp = TP(0, tp_size, fabric=Fabric.emulated(tp_size, cfg["text_config"]["hidden_size"], "cuda")[0], emulated=True)
    eng = PaliGemmaEngine(cfg, gpu_weights(cfg, torch.bfloat16), tp=tp)
    d, L, st, dt = eng.dims, cabi.lib(), cabi.stream(), eng.dt
    nq, F_l, V_l = eng.nq_l, eng.F_l, eng.V_l
    kv = eng.new_kv(B)
    kv.reserve(T + 600)
    kv.length = T
    kv.kv_len.fill_(T)
    ds = eng.decode_state(B)
    ds.bind(kv, torch.full((B,), 5, device="cuda"), T + 1)
    ds.x.normal_(); ds.x2.normal_(); ds.att.normal_(); ds.g.normal_(); ds.q.normal_()
    nl = len(eng.t_layers)
    e = 2
    res = {}
    if B >= eng.batched_min:   # batched step only (skinny-GEMM path)
        ds.run_steps(kv, 1)
        g = next(iter(ds.graphs.values()))
        med, mn = timeit(lambda: [g.replay() for _ in range(10)], 10)
        res = {"batch": B, "context": T, "graph_step_us": {"median": round(med, 1), "min": round(mn, 1)},
               "tokens_per_s": round(B / (med * 1e-6), 1), "step_bytes": eng.weight_bytes_per_decode_step()}
        res["step_frac_of_6555"] = round((res["step_bytes"] + B * d.L * 2 * d.hd * 2 * T) / (med * 1e-6) / 1e9 / 6555.2, 4)
        print(json.dumps(res))
        return

    def rec(name, fn, nbytes):
        med, mn = timeit(fn, nl)
        res[name] = {"us": round(med, 2), "min_us": round(mn, 2), "GBps": round(nbytes / med / 1e3, 1)}

    def qkv():
        for li, w in enumerate(eng.t_layers):
            L.pg_decode_qkv(ds.q.data_ptr(), ds.x.data_ptr(), w["ln1"].data_ptr(), w["qkv"].data_ptr(), eng.inv_freq.data_ptr(),
                            ds.pos.data_ptr(), eng.k_pool[li].data_ptr(), eng.v_pool[li].data_ptr(), kv.page_table.data_ptr(),
                            kv.max_pages, eng.page_size, kv.kv_len.data_ptr(), B, d.D, nq, d.nkv, d.hd, d.eps, d.max_pos, None, None, dt, st)
    rec("qkv", qkv, e * (nq + 2 * d.nkv) * d.hd * d.D)

    def attn():
        for li, w in enumerate(eng.t_layers):
            L.pg_decode_attention(ds.att.data_ptr(), ds.q.data_ptr(), eng.k_pool[li].data_ptr(), eng.v_pool[li].data_ptr(),
                                  kv.page_table.data_ptr(), kv.max_pages, eng.page_size, kv.kv_len.data_ptr(), 1, B, nq, d.nkv,
                                  d.hd, 16.0, None, None, 32, dt, st)
    rec("attention", attn, e * B * 2 * (T + 1) * d.nkv * d.hd)

    def oproj():
        for li, w in enumerate(eng.t_layers):
            L.pg_gemv_res(ds.x2.data_ptr(), ds.att.data_ptr(), w["o"].data_ptr(), ds.x.data_ptr(), B, d.D, nq * d.hd, None, dt, st)
    rec("o_proj", oproj, e * d.D * nq * d.hd)

    def gateup():
        for li, w in enumerate(eng.t_layers):
            L.pg_decode_gateup(ds.g.data_ptr(), ds.x2.data_ptr(), w["ln2"].data_ptr(), w["gu"].data_ptr(), B, d.D, F_l, d.eps, None, None, dt, st)
    rec("gateup", gateup, e * 2 * F_l * d.D)

    def down():
        for li, w in enumerate(eng.t_layers):
            L.pg_gemv_res(ds.x.data_ptr(), ds.g.data_ptr(), w["down"].data_ptr(), ds.x2.data_ptr(), B, d.D, F_l, None, dt, st)
    rec("down", down, e * F_l * d.D)

    def lmhead():
        L.pg_decode_lmhead(ds.local_logits.data_ptr(), ds.x.data_ptr(), eng.final_norm.data_ptr(), eng.lm_head.data_ptr(), B, d.D, V_l,
                           d.eps, ds.keys.data_ptr(), None, None, dt, st)
    med, mn = timeit(lmhead, 1)
    res["lm_head"] = {"us": round(med, 2), "min_us": round(mn, 2), "GBps": round(e * V_l * d.D / med / 1e3, 1)}

    per_layer = sum(res[k]["us"] for k in ("qkv", "attention", "o_proj", "gateup", "down"))
    res["sum_isolated_us"] = round(nl * per_layer + res["lm_head"]["us"], 1)
    if tp_size > 1:      # an emulated rank cannot run a step alone (its consumers would wait for the other ranks)
        from pg_b200._cabi import exref
        fab = eng.fabric
        L.pg_tp_begin_step(fab.epoch.data_ptr(), st)
        ex = fab.x(1, 2 * d.L + 2)

        def oproj_push():       # producer variant of o_proj: partials stored into all ranks' exchange buffers
            for li, w in enumerate(eng.t_layers):
                L.pg_gemv_res(None, ds.att.data_ptr(), w["o"].data_ptr(), None, B, d.D, nq * d.hd, exref(ex), dt, st)
        rec("o_proj_push", oproj_push, e * d.D * nq * d.hd)
        res["tp"] = tp_size
        print(json.dumps(res))
        return
    # whole step through the graph
    ds.run_steps(kv, 1)
    g = next(iter(ds.graphs.values()))
    med, mn = timeit(lambda: [g.replay() for _ in range(20)], 20)
    res["graph_step_us"] = {"median": round(med, 1), "min": round(mn, 1)}
    res["step_bytes"] = eng.weight_bytes_per_decode_step()
    res["step_frac_of_6555"] = round(res["step_bytes"] / (med * 1e-6) / 1e9 / 6555.2, 4)
    res["env"] = {k: v for k, v in os.environ.items() if k.startswith("PG_") or k.startswith("SWEEP_")}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
