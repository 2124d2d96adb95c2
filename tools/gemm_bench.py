#!/usr/bin/env python
"""Time pg_gemm (tcgen05 paths) on the SigLIP batch-64 / prefill shapes; TFLOP/s per shape.
Tunables come from the environment (PG_GEMM_2CTA, PG_GEMM_MCAST, PG_GEMM_BN ...)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200"))
import torch  # noqa: E402
from pg_b200 import _cabi as cabi  # noqa: E402

SHAPES = [  # name, M, N, K, epilogue
    ("qkv", 16384, 3840, 1152, cabi.EPI_BIAS),
    ("o_proj", 16384, 1152, 1152, cabi.EPI_BIAS_RES),
    ("fc1", 16384, 4304, 1152, cabi.EPI_BIAS_GELU),
    ("fc2", 16384, 1152, 4304, cabi.EPI_BIAS_RES),
    ("proj", 16384, 2048, 1152, cabi.EPI_BIAS),
    ("lm_head_260", 260, 257216, 2048, cabi.EPI_NONE),
]


def main():
    dt = torch.bfloat16
    L, st = cabi.lib(), cabi.stream()
    res = {"env": {k: v for k, v in os.environ.items() if k.startswith("PG_")}}
    for name, M, N, K, epi in SHAPES:
        a = torch.randn(M, K, device="cuda", dtype=dt)
        w = torch.randn(N, K, device="cuda", dtype=dt) * 0.05
        b = torch.randn(N, device="cuda", dtype=dt)
        r = torch.randn(M, N, device="cuda", dtype=dt) if epi == cabi.EPI_BIAS_RES else None
        out = torch.empty(M, N, device="cuda", dtype=dt)

        def run():
            cabi.check(L.pg_gemm(out.data_ptr(), a.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr() if r is not None else None,
                                 M, N, K, K, K, N, N, 0, epi, 0, 2, cabi.DTYPE_CODE[dt], st), name)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 10)
        ms = sorted(ts)[len(ts) // 2]
        res[name] = {"us": round(ms * 1e3, 1), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
        del a, w, b, r, out
    print(json.dumps(res))


if __name__ == "__main__":
    main()
