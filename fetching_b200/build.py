"""Build fetching_b200/lib/libpfetch.so (the C-ABI engine) with nvcc for sm_100a, in-tree.

    python -m fetching_b200.build [--force] [--verbose]

The shared object travels to the GPU box with the repository snapshot; nothing is JIT-compiled
at run time.  There is no CPU build of the engine: the kernels are sm_100a only.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libpfetch.so")

CU_SOURCES = ["pf_pointwise.cu", "pf_resident.cu", "pf_matvec.cu", "pf_matvec_tc.cu", "pf_engine.cu", "pf_probe.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-I" + os.path.join(ROOT, "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine cannot be built (there is no CPU fallback)")


def _sources():
    out = []
    for d, _, files in os.walk(CSRC):
        out += [os.path.join(d, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp"))]
    out.append(os.path.join(ROOT, "include", "pfetch.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False, out: str = LIB, defines=()) -> str:
    """`out` / `defines` build an experimental variant next to the default library (PFETCH_LIB selects it)."""
    if out == LIB and not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj" if out == LIB else "obj_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    # the distro g++ is the host compiler nvcc 12.9 supports; $CXX in this image may point elsewhere
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    procs = []
    objs = []
    for src in CU_SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *ccbin, *NVCC_FLAGS, *["-D" + d for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(log)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    host_cpp = sorted(os.path.join(CSRC, "host", f) for f in os.listdir(os.path.join(CSRC, "host"))
                      if f.endswith(".cpp") and not f.endswith("_main.cpp"))
    link = [nvcc, *ccbin, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17",
            "-Xcompiler", "-fPIC,-O2", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
            "-o", out, *objs, *host_cpp, "-lcuda", "-ldl"]
    if verbose:
        print(" ".join(link), flush=True)
    subprocess.run(link, check=True, env=env)
    if out == LIB:  # the command-line front end (tree.cc equivalent), linked against the library next to it
        cli = [ccbin[1] if ccbin else "g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
               os.path.join(CSRC, "host", "pfetch_main.cpp"), "-o", os.path.join(LIBDIR, "pfetch"),
               "-L" + LIBDIR, "-lpfetch", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cli), flush=True)
        subprocess.run(cli, check=True, env=env)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                out=os.path.join(LIBDIR, outs[0]) if outs else LIB, defines=defs))
