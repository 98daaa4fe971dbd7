"""Build ``lib/libgss.so`` (hand-written sm_100a kernels + the C ABI) with nvcc.

``python -m gan_sass_tf_b200.build`` - in-tree, so the library travels with the
source tree to the GPU box.  nvcc cross-compiles without a GPU.

``csrc/gss_api.cu`` is compiled five times with ``-DGSS_PART=0..4`` (C ABI + element-wise
kernels / streaming kernels N = 512 / N = 256 / team kernels N <= 1024 / N >= 2048), in
parallel, and the objects are linked into one shared library.

Two flavours come out of one build: ``lib/libgss.so`` (the product: no process-wide switches, no
measured-slower kernel variants) and ``lib/libgss_experimental.so`` (parts 0, 1 and 3 recompiled with
``-DGSS_EXPERIMENTAL``: ``gss_set_path`` / ``gss_set_synth_variant``, the role-split and tensor-memory
synthesis variants and the team kernels at N = 256 / 512 that the cross-check tests and the tuning
tools use; parts 2 and 4 are shared).  Eight compile jobs in parallel, about three minutes on 8 cores.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libgss.so")
OUT_EXPERIMENTAL = os.path.join(HERE, "lib", "libgss_experimental.so")
EXPERIMENTAL_PARTS = (0, 1, 3)
SOURCE = "gss_api.cu"
PARTS = (0, 1, 2, 3, 4)
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "gss_api.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    # inline variables and the statics of template functions are STB_GNU_UNIQUE by default: the dynamic linker would then
    # share them between libgss.so and libgss_experimental.so inside one process even under RTLD_LOCAL (e.g. the
    # "lane tables filled" flags, while each library owns its own __device__ tables)
    "-Xcompiler", "-fno-gnu-unique",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(OUT) or not os.path.exists(OUT_EXPERIMENTAL):
        return True
    t = min(os.path.getmtime(OUT), os.path.getmtime(OUT_EXPERIMENTAL))
    deps = [os.path.join(CSRC, f) for f in [SOURCE] + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = _nvcc()
    src = os.path.join(CSRC, SOURCE)
    extra = ["-Xptxas", "-v"] if verbose else []
    obj_dir = tempfile.mkdtemp(prefix="gss_obj_")          # objects stay out of the tree (only the .so travels)
    try:
        objs = [os.path.join(obj_dir, f"gss_part{k}.o") for k in PARTS]
        xobjs = {k: os.path.join(obj_dir, f"gss_xpart{k}.o") for k in EXPERIMENTAL_PARTS}
        cmds = [[nvcc] + NVCC_FLAGS + extra + [f"-DGSS_PART={k}", "-c", src, "-o", o] for k, o in zip(PARTS, objs)]
        cmds += [[nvcc] + NVCC_FLAGS + extra + ["-DGSS_EXPERIMENTAL", f"-DGSS_PART={k}", "-c", src, "-o", o] for k, o in xobjs.items()]
        # the slowest parts first (1 and its experimental twin carry the most kernel instances)
        cmds.sort(key=lambda c: 0 if "-DGSS_PART=1" in c else 1)
        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as ex:
            list(ex.map(lambda c: _run(c, verbose), cmds))
        link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"]
        _run(link + [OUT] + objs, verbose)
        _run(link + [OUT_EXPERIMENTAL] + [xobjs.get(k, objs[i]) for i, k in enumerate(PARTS)], verbose)
    finally:
        shutil.rmtree(obj_dir, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
