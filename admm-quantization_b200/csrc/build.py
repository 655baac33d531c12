"""Build libadmmq.so (sm_100a) in-tree:  python admm-quantization_b200/csrc/build.py

nvcc cross-compiles without a GPU.  The shared library lands in
admm-quantization_b200/lib/libadmmq.so (git-ignored; it travels to the GPU box with gpurun).
"""
import os
import shutil
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(PKG, "build")
SOURCES = ["api.cu", "project.cu", "contract.cu", "admm_loop.cu", "mttkrp_tc.cu", "tc_gemm.cu", "factorize.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--fmad=true", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "admmq.h"))
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [os.path.join(HERE, src)] + headers):
            cmd = [nvcc] + FLAGS + ["-c", os.path.join(HERE, src), "-o", obj]
            procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, cmd, pr in procs:
        out, _ = pr.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if pr.returncode != 0:
            sys.stderr.write(log[-1])
            raise RuntimeError(f"nvcc failed on {src}")
    lib = os.path.join(OUT_DIR, "libadmmq.so")
    if procs or not os.path.exists(lib):
        cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(f"$ {' '.join(cmd)}\n{r.stdout}")
        if r.returncode != 0:
            sys.stderr.write(log[-1])
            raise RuntimeError("link failed")
    with open(os.path.join(OBJ_DIR, "build.log"), "a") as f:
        f.write(f"==== {time.ctime()}\n" + "\n".join(log) + "\n")
    if verbose:
        print("\n".join(log))
    return lib


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
