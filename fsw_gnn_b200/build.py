"""Build libfsw_embedding.so in-tree with nvcc for sm_100a (the new `build_fsw_embedding`,
reference: build_fsw_embedding:16-24 which stops at sm_86).

    python -m fsw_gnn_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfsw_embedding.so")
SOURCES = ["fsw_api.cu", "fsw_umma.cu", "fsw_prep.cu", "fsw_gemm.cu", "fsw_segcumsum.cu", "fsw_embed.cu", "fsw_embed_medium.cu", "fsw_embed_small.cu", "fsw_embed_packed.cu", "fsw_wgrad.cu", "fsw_cloud.cu"]
NVCC_FLAGS = ["-std=c++17", "--expt-relaxed-constexpr", "-O3", "-lineinfo",
              "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]


def _nvcc():
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cand = os.path.join(cuda_home, "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "fsw_embedding.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    if not force and not _stale():
        if verbose:
            print("libfsw_embedding.so is up to date")
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        srcp = os.path.join(CSRC, src)
        if (not force) and os.path.exists(obj) and os.path.getmtime(obj) > max(
                os.path.getmtime(srcp), os.path.getmtime(os.path.join(CSRC, "fsw_common.cuh")),
                os.path.getmtime(os.path.join(CSRC, "fsw_sortnet.cuh")),
                os.path.getmtime(os.path.join(HERE, "..", "include", "fsw_embedding.h"))):
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", srcp, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [_nvcc(), "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
