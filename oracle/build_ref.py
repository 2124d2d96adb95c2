#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: the reference's own Python files for the hot path, taken UNMODIFIED from
/root/reference (read-only, present in the build container only).  TEST INFRASTRUCTURE ONLY.

The reference is pure Python, so there is nothing to compile: "building" it means staging the five files the path
consists of where the GPU box can import them (`oracle/_ref/` is git-ignored -- no reference source enters the
history -- but travels with the gpurun snapshot, like the built .so):

  modeling_gemma.py, modeling_siglip.py      the model the CPU baseline / `bench.py --impl reference` times
  processing_paligemma.py                    its processor
  ablation_study_fixed.py, inference.py      the two drivers tests/test_gpu_reference_drivers.py runs against the drop-ins

    python oracle/build_ref.py        (also run by __graft_entry__.build() when /root/reference exists)
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ("modeling_gemma.py", "modeling_siglip.py", "processing_paligemma.py", "ablation_study_fixed.py", "inference.py")


def build(src: str = "/root/reference") -> bool:
    if not os.path.isdir(src):
        return False
    os.makedirs(OUT, exist_ok=True)
    lines = []
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(OUT, f))
        with open(os.path.join(OUT, f), "rb") as fh:
            lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {f}")
    with open(os.path.join(OUT, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, f)) for f in FILES)


def import_reference():
    """(modeling_gemma, modeling_siglip) of the REFERENCE, imported from oracle/_ref under their own names without
    leaving them in sys.modules (the drop-in modules of this repo carry the same names)."""
    saved = {m: sys.modules.pop(m, None) for m in ("modeling_gemma", "modeling_siglip")}
    sys.path.insert(0, OUT)
    try:
        import modeling_gemma as ref_gemma      # noqa: the reference's own module
        import modeling_siglip as ref_siglip    # noqa
    finally:
        sys.path.remove(OUT)
        for m, mod in saved.items():
            sys.modules.pop(m, None)
            if mod is not None:
                sys.modules[m] = mod
    return ref_gemma, ref_siglip


if __name__ == "__main__":
    ok = build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("oracle/_ref staged" if ok else "no reference tree: nothing staged")
