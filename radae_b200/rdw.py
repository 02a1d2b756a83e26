"""Weight containers for the RADE core codec (host-side, numpy only).

Two formats are understood:

* **DNNw** — the reference's run-time blob (`bin/model19_check3.bin`): a sequence of 64-byte records
  `{"DNNw", int version, int type, int size, int block_size, char name[44]}` each followed by the payload
  padded to 64 bytes (reference: src/write_rade_weights.c:51-74).  int8 matrices are stored as
  8(out) x 4(in) blocks in (out/8, in/4) block order (weight-exchange/wexchange/c_export/common.py:59-67);
  GRU input matrices additionally carry a `[count, in_pos...]` block index list per 8 outputs (:156-170).

* **RDW** — this repo's own container (what `libradae_b200.so` embeds and what the CUDA weight upload
  consumes): plain row-major matrices, no blocking, no index lists:
      `<layer>.w8`    int8  [out][in]      (int8 layers)
      `<layer>.scale` f32   [out]          per-output scale  (= exporter scale/127, common.py:267)
      `<layer>.bias`  f32   [out]
      `<layer>.wf`    f32   [in][out]      (the four float layers; [in][out] as the reference stores them)
  Layout: 64-byte file header, n x 80-byte entries, 64-byte aligned payloads.
"""
import struct
import numpy as np

RDW_MAGIC = b"RADEB200"
RDW_VERSION = 1
DT_F32, DT_I8, DT_I32 = 0, 1, 2
_NP = {DT_F32: np.float32, DT_I8: np.int8, DT_I32: np.int32}

# (layer, nb_inputs, nb_outputs, kind) — shapes as in src/rade_enc_data.c:227866-227882 and
# src/rade_dec_data.c:222152-222173 (model19_check3: input_dim = output_dim = 84)
ENC_LAYERS = [("enc_dense1", 84, 64, "f32")] + \
    [l for i, k in enumerate((64, 224, 384, 544, 704), 1)
     for l in ((f"enc_gru{i}_input", k, 192, "i8s"), (f"enc_gru{i}_recurrent", 64, 192, "i8"))] + \
    [(f"enc_conv{i}", k, 96, "i8") for i, k in enumerate((256, 576, 896, 1216, 1536), 1)] + \
    [("enc_zdense", 864, 80, "f32")]
DEC_LAYERS = [("dec_dense1", 80, 96, "f32")] + \
    [l for i, k in enumerate((96, 224, 352, 480, 608), 1)
     for l in ((f"dec_gru{i}_input", k, 288, "i8s"), (f"dec_gru{i}_recurrent", 96, 288, "i8"))] + \
    [(f"dec_glu{i}", 96, 96, "i8") for i in range(1, 6)] + \
    [(f"dec_conv{i}", k, 32, "i8") for i, k in enumerate((384, 640, 896, 1152, 1408), 1)] + \
    [("dec_output", 736, 84, "f32")]
ALL_LAYERS = ENC_LAYERS + DEC_LAYERS


def parse_dnnw(buf):
    """DNNw blob -> {name: (type, ndarray)}; type 0 float, 1 int, 3 int8."""
    out = {}
    off = 0
    while off < len(buf):
        head, version, typ, size, block_size = struct.unpack_from("<4siiii", buf, off)
        if head != b"DNNw" or version != 0:
            raise ValueError(f"bad DNNw record at byte {off}")
        name = buf[off + 20:off + 64].split(b"\0")[0].decode()
        dt = {0: np.float32, 1: np.int32, 3: np.int8}[typ]
        out[name] = (typ, np.frombuffer(buf, dtype=dt, count=size // np.dtype(dt).itemsize, offset=off + 64).copy())
        off += 64 + block_size
    return out


def unblock_int8(w_blocks, nb_in, nb_out, idx=None):
    """8x4-blocked int8 (optionally block-indexed) -> dense [out][in] int8."""
    W = np.zeros((nb_out, nb_in), np.int8)
    p = 0
    ip = 0
    for ob in range(nb_out // 8):
        if idx is None:
            positions = range(0, nb_in, 4)
        else:
            n = int(idx[ip]); ip += 1
            positions = [int(v) for v in idx[ip:ip + n]]; ip += n
        for pos in positions:
            W[ob * 8:ob * 8 + 8, pos:pos + 4] = w_blocks[p:p + 32].reshape(8, 4)
            p += 32
    if p != w_blocks.size:
        raise ValueError("int8 block count mismatch")
    return W


def dnnw_to_arrays(buf):
    """DNNw blob -> flat {name: ndarray} in RDW naming (row-major, unblocked)."""
    raw = parse_dnnw(buf)
    arrays = {}
    for name, nin, nout, kind in ALL_LAYERS:
        arrays[f"{name}.bias"] = raw[f"{name}_bias"][1].astype(np.float32)
        # models without the auxiliary symbol (the reference's bin/model05.bin): 80 encoder inputs / decoder outputs
        if name == "enc_dense1" and raw[f"{name}_weights_float"][1].size == 80 * nout:
            nin = 80
        if name == "dec_output" and arrays[f"{name}.bias"].size == 80:
            nout = 80
        if kind == "f32":
            wf = raw[f"{name}_weights_float"][1]
            arrays[f"{name}.wf"] = wf.reshape(nin, nout).astype(np.float32)
        else:
            idx = raw[f"{name}_weights_idx"][1] if kind == "i8s" else None
            arrays[f"{name}.w8"] = unblock_int8(raw[f"{name}_weights_int8"][1], nin, nout, idx)
            arrays[f"{name}.scale"] = raw[f"{name}_scale"][1].astype(np.float32)
    return arrays


def write_rdw(path, arrays, model_name="model19_check3"):
    names = list(arrays.keys())
    n = len(names)
    off = 64 + 80 * n
    off = (off + 63) // 64 * 64
    entries, payload = [], []
    for nm in names:
        a = np.ascontiguousarray(arrays[nm])
        dt = {np.dtype(np.float32): DT_F32, np.dtype(np.int8): DT_I8, np.dtype(np.int32): DT_I32}[a.dtype]
        rows, cols = (a.shape[0], a.shape[1]) if a.ndim == 2 else (1, a.shape[0])
        nbytes = a.nbytes
        entries.append(struct.pack("<48sIIIIQQ", nm.encode(), dt, rows, cols, 0, off, nbytes))
        payload.append((off, a.tobytes()))
        off = (off + nbytes + 63) // 64 * 64
    with open(path, "wb") as f:
        f.write(struct.pack("<8sIIQ40s", RDW_MAGIC, RDW_VERSION, n, off, model_name.encode()))
        for e in entries:
            f.write(e)
        for o, b in payload:
            f.seek(o)
            f.write(b)
        f.truncate(off)


def read_rdw(path_or_bytes):
    buf = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    magic, version, n, total, _ = struct.unpack_from("<8sIIQ40s", buf, 0)
    if magic != RDW_MAGIC or version != RDW_VERSION:
        raise ValueError("not an RDW v1 file")
    arrays = {}
    for i in range(n):
        nm, dt, rows, cols, _, off, nbytes = struct.unpack_from("<48sIIIIQQ", buf, 64 + 80 * i)
        a = np.frombuffer(buf, dtype=_NP[dt], count=nbytes // np.dtype(_NP[dt]).itemsize, offset=off)
        arrays[nm.split(b"\0")[0].decode()] = a.reshape(rows, cols) if rows > 1 else a.reshape(cols)
    return arrays


def default_weights_path():
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights", "model19_check3.rdw")


def model05_weights_path():
    """the reference's second shipped core codec (bin/model05.bin: no aux symbol, bottleneck 1), converted by tools/export_weights.py"""
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights", "model05.rdw")
