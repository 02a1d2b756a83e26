"""Build the oracle's native pieces (TEST INFRASTRUCTURE — not the product).

  oracle/_ref/librade_ref_int8.so   reference C core codec (its own rade_enc.c, rade_dec.c,
  oracle/_ref/librade_ref_f32.so    rade_*_data.c compiled where they lie under /root/reference/src)
                                    + oracle/nnet_shim; int8 = -DDISABLE_DEBUG_FLOAT (what ships),
                                    f32 = float debug weights (validates layouts against PyTorch)
  oracle/_ref/libcore_oracle.so     our own C restatement of the core codec (oracle/core_oracle.c)

Only built here when /root/reference exists; the GPU box uses the prebuilt .so files that travel
with the snapshot (oracle/_ref is git-ignored, not gpurun-ignored).  The reference's own build
system (cmake + network fetch of opus) is NOT run: opus is un-vendored, see DESIGN.md.
"""
import os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
OUT = os.path.join(HERE, "_ref")
CFLAGS = ["-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", "-fvisibility=hidden", "-Wall",
          "-Wno-unused-variable", "-Wno-unused-function"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd):
    print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)


def build_ref(force=False):
    """Compile the reference C core codec + shim.  Returns True if available afterwards."""
    os.makedirs(OUT, exist_ok=True)
    shim = [os.path.join(HERE, "nnet_shim", "nnet_shim.c"), os.path.join(HERE, "ref_core_api.c")]
    have_ref = os.path.isdir(REF_SRC)
    ok = True
    for variant, defs in (("int8", ["-DDISABLE_DEBUG_FLOAT"]), ("f32", [])):
        target = os.path.join(OUT, f"librade_ref_{variant}.so")
        if not have_ref:
            ok = ok and os.path.exists(target)
            continue
        ref = [os.path.join(REF_SRC, f) for f in ("rade_enc.c", "rade_dec.c", "rade_enc_data.c", "rade_dec_data.c")]
        if force or _stale(target, shim + [os.path.join(HERE, "nnet_shim", "nnet.h")]):
            _run(["gcc"] + CFLAGS + defs + ["-I", os.path.join(HERE, "nnet_shim"), "-I", REF_SRC]
                 + ref + shim + ["-lm", "-o", target])
    return ok


def build_port(force=False):
    os.makedirs(OUT, exist_ok=True)
    target = os.path.join(OUT, "libcore_oracle.so")
    src = [os.path.join(HERE, "core_oracle.c")]
    if force or _stale(target, src):
        _run(["gcc"] + CFLAGS + src + ["-lm", "-o", target])
    return True


def build_all(force=False):
    build_port(force)
    return build_ref(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
