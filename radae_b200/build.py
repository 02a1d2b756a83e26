"""Build libradae_b200.so in-tree with nvcc for sm_100a (no torch involved; plain CUDA runtime).

    python -m radae_b200.build            # or: from radae_b200.build import build; build()

Output: radae_b200/lib/libradae_b200.so  (git-ignored, travels to the GPU box with the snapshot).
"""
import os, subprocess, sys, hashlib

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libradae_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
WEIGHTS = os.path.join(HERE, "weights", "model19_check3.rdw")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function,-fopenmp", "-DIS_BUILDING_RADE_API=1",
              "-I", INCLUDE, "-I", CSRC]


def _sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(f for f in os.listdir(CSRC) if f.endswith(".cpp"))
    return cu, cpp


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [WEIGHTS, os.path.join(INCLUDE, "rade_api.h"),
                                                                os.path.join(INCLUDE, "rade_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    cu, cpp = _sources()
    objs = []
    procs = []
    for f in cu + cpp:
        o = os.path.join(objdir, f + ".o")
        objs.append(o)
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, f), "-o", o]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    wo = os.path.join(objdir, "embed_weights.o")
    subprocess.run(["gcc", "-c", os.path.join(CSRC, "embed_weights.S"), f'-DRADE_WEIGHTS_FILE="{WEIGHTS}"', "-o", wo], check=True)
    objs.append(wo)
    failed = False
    for f, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {f}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs +
                   ["-Xlinker", "--no-undefined", "-lcudart", "-lgomp"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
