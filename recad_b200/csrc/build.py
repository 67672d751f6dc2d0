"""Build librecad_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m recad_b200.csrc.build [--force] [--verbose]

The shared library travels to the GPU box with the gpurun snapshot; it is
git-ignored.  Only sm_100a SASS is embedded: there is no PTX fallback and no
other architecture.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "librecad_b200.so")
SOURCES = ["api.cu", "scan.cu", "sort.cu", "csr.cu", "spmm.cu", "bpr.cu", "shard.cu", "wmf.cu", "aush.cu", "mf.cu", "ncf.cu", "eval.cu", "eval_tc.cu", "gemm_tc.cu", "sampler.cpp"]
HEADERS = ["common.cuh", "rank_epilogue.cuh", "tc_common.cuh", os.path.join(ROOT, "include", "recad_b200.h")]
# host-only translation units built by g++ with ISA flags of their own (each is entered only after a run-time CPU check)
CXX_SOURCES = {"sampler_avx512.cpp": ["-mavx512f", "-mbmi", "-mlzcnt"]}
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
]


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(HERE, "_build", os.path.splitext(src)[0] + ".o")
    path = os.path.join(HERE, src)
    deps = [path] + [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    if not _stale(obj, deps):
        return obj
    if src in CXX_SOURCES:
        cmd = [CXX, "-O3", "-std=c++17", "-fPIC", "-Wall"] + CXX_SOURCES[src] + ["-c", path, "-o", obj]
    else:
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose or r.stderr.strip():
        sys.stderr.write(f"--- {src}\n{r.stderr}")
    return obj


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    srcs = [s for s in SOURCES + list(CXX_SOURCES) if os.path.exists(os.path.join(HERE, s))]
    if force:
        for f in os.listdir(os.path.join(HERE, "_build")):
            os.remove(os.path.join(HERE, "_build", f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
