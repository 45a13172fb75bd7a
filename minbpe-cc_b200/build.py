"""Build recipe for libminbpe_b200.so and the minbpe-cc CLI (in-tree, sm_100a only).

    python minbpe-cc_b200/build.py [--force]

nvcc cross-compiles for sm_100a without a GPU. Outputs: minbpe-cc_b200/lib/libminbpe_b200.so and
minbpe-cc_b200/bin/minbpe-cc (git-ignored; they travel to the GPU box with gpurun).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "lib", "libminbpe_b200.so")
CLI = os.path.join(HERE, "bin", "minbpe-cc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CU = ["capi.cu", "train_cuda.cu", "encode.cu", "pretok.cu"]
CPP = ["chunker.cpp", "tokenizer.cpp"]
EXTRA_DEFS = [d for d in os.environ.get("MBPE_DEFS", "").split() if d]  # e.g. MBPE_DEFS="-DMBPE_PROFILE_HITS"
NVCC_FLAGS = EXTRA_DEFS + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC", "-cudart", "static"]
CXX_FLAGS = ["-std=c++23", "-O3", "-fPIC", "-Wall", "-pthread"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + cmd[0])
    return r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "lib"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "bin"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
    headers += [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".hpp", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "minbpe_b200.h"))
    objs = []
    for f in CU:
        src, obj = os.path.join(CSRC, f), os.path.join(HERE, "build", f + ".o")
        if force or _newer(obj, [src] + headers):
            out = _run([NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
            if verbose:
                print(out)
        objs.append(obj)
    cxx = os.environ.get("CXX", "g++")
    for f in CPP:
        src, obj = os.path.join(HOST, f), os.path.join(HERE, "build", f + ".o")
        if force or _newer(obj, [src] + headers):
            _run([cxx] + CXX_FLAGS + ["-c", src, "-o", obj])
        objs.append(obj)
    if force or _newer(LIB, objs):
        _run([NVCC, "-shared", "-cudart", "static", "-o", LIB] + objs + ["-Xlinker", "-l:libpcre2-8.so.0", "-lpthread"])
    cli_src = os.path.join(HOST, "cli.cpp")
    if force or _newer(CLI, [cli_src, LIB] + headers):
        _run([cxx] + CXX_FLAGS + [cli_src, "-o", CLI, "-L" + os.path.join(HERE, "lib"), "-lminbpe_b200",
                                  "-Wl,-rpath,$ORIGIN/../lib", "-l:libpcre2-8.so.0"])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
