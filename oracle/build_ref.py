"""Build recipe for ``oracle/_ref/`` — the UNMODIFIED reference, compiled (test / baseline infrastructure).

The reference is pure Python, so "building" it means byte-compiling its modules FROM THE SOURCES WHERE THEY LIE under
``/root/reference`` into sourceless compiled modules (``.refc``: the ``.pyc`` format under an extension snapshot tools do not drop) that mirror its package layout:

    python oracle/build_ref.py            # also run by __graft_entry__.build() when /root/reference is present

Outputs go ONLY into ``oracle/_ref/`` (git-ignored so no reference source or artefact enters the history, but not
gpurun-ignored, so the compiled tree travels to the GPU box like the repo's own ``.so``).  No reference source file is
copied.  The GPU box runs the same image (same CPython magic number), so ``oracle/_ref_import.install()`` can import the
compiled reference there: ``bench.py --impl reference`` (the reference's own CPU path, ``kind: "reference"``),
``bench.py``'s ``gpu_eager_reference`` leg and the reference-on-GPU deviation tests use it.  Product code never does.
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = os.environ.get("DSDIFF_REFERENCE_ROOT", "/root/reference")

# packages the sampling hot path and its callers import
PACKAGES = ["ldm", "UNet_DS_Diff", "training_project", "trainers", "Disc_diff/guided_diffusion", "loss_function"]
# compiled-module extension: NOT ".pyc" — snapshot tools (gpurun among them) drop *.pyc as cache files, so the compiled
# reference would silently stay behind; oracle/_ref_import.py registers a sourceless loader for this extension
EXT = ".refc"
# namespace parents that hold no needed code of their own
PARENTS = ["Disc_diff"]


def build(src: str = SRC, out: str = OUT) -> int:
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} not found: oracle/_ref can only be (re)built in the authoring container")
    if os.path.isdir(out):
        shutil.rmtree(out)
    n = 0
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(src, pkg)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            rel = os.path.relpath(root, src)
            for f in files:
                if not f.endswith(".py"):
                    continue
                dst = os.path.join(out, rel, f[:-3] + EXT)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                try:
                    # dfile: the path tracebacks show (the reference's own), checked-hash pyc without source validation
                    py_compile.compile(os.path.join(root, f), cfile=dst, dfile=os.path.join("/root/reference", rel, f),
                                       doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
                    n += 1
                except py_compile.PyCompileError as e:      # a few unrelated reference files do not parse on 3.12
                    print(f"build_ref: skipped {rel}/{f}: {e.msg.strip().splitlines()[-1]}", file=sys.stderr)
    for parent in PARENTS:
        os.makedirs(os.path.join(out, parent), exist_ok=True)
    with open(os.path.join(out, "BUILD_INFO.txt"), "w") as fh:
        fh.write(f"byte-compiled from {src} by oracle/build_ref.py with CPython {sys.version.split()[0]}; {n} modules\n")
    return n


if __name__ == "__main__":
    print(f"oracle/_ref: {build()} modules compiled")
