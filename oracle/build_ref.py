"""Recipe for `oracle/_ref/`: the UNMODIFIED reference module of the hot path, placed where it can travel to the GPU box.

    python oracle/build_ref.py            (build container only; needs /root/reference)

`/root/reference/CookieTTS/_4_mtw/waveglow/glow.py` is a single torch-only module (its own imports: torch,
torch.autograd.Variable, torch.nn.functional).  The reference has no build system for it, so "compiling" it is a
byte-for-byte copy into `oracle/_ref/glow.py` plus a SHA-256 of the source beside it.  `oracle/_ref/` is listed in
.gitignore (nothing of the reference enters the history) but not in .gpurunignore, so `bench.py --impl reference`
and the `cpu_baseline` leg can time the reference's own `WaveGlow.infer` on the GPU box's host cores.
`__graft_entry__.build()` runs this when /root/reference is present.  Test infrastructure: only bench.py's CPU legs
and tests/ may import anything under oracle/.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/CookieTTS/_4_mtw/waveglow/glow.py"
DST_DIR = os.path.join(HERE, "_ref")


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref/glow.py is in place (copied now, or already there)."""
    dst = os.path.join(DST_DIR, "glow.py")
    if not os.path.exists(SRC):
        return os.path.exists(dst)
    os.makedirs(DST_DIR, exist_ok=True)
    shutil.copyfile(SRC, dst)
    digest = hashlib.sha256(open(SRC, "rb").read()).hexdigest()
    with open(os.path.join(DST_DIR, "SOURCE.txt"), "w") as f:
        f.write(f"{SRC}\nsha256 {digest}\ncopied unmodified by oracle/build_ref.py\n")
    if verbose:
        print(f"oracle/_ref/glow.py <- {SRC} (sha256 {digest[:16]}...)")
    return True


def load():
    """Import oracle/_ref/glow.py as a module (None if the recipe has not been run)."""
    path = os.path.join(DST_DIR, "glow.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("cookietts_ref_glow", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
