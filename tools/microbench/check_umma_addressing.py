"""CPU check of the address arithmetic in the tcgen05 prototypes (umma_gru_layer.cu, umma_conv_layer.cu, umma_gru_chain.cu,
umma_tf32_refresh.cu) against the canonical no-swizzle K-major operand layout as documented in CUTLASS
(cute/atom/mma_traits_sm100.hpp: ((8,m),(T,2)):((1T,SBO),(1,LBO)) in 16-byte units):

    byte address of (row r, k byte b) of an operand = start + (r // 8) * SBO + (b // 16) * LBO + (r % 8) * 16 + b % 16

The emulated MMA reads a 128-row A tile and an N-row B tile through exactly the (start, LBO, SBO) triples the CUDA sources
build, accumulates D[lane][column], and the result is compared with a plain integer / float GEMM of the logical matrices.
It cannot tell whether the hardware agrees with that layout (the probes on the GPU do), but it catches slips in the
prototypes' own offsets — tile starts of the overlapping gate tiles, the 256-byte k-block advance, chunk strides.
Run: python tools/microbench/check_umma_addressing.py"""
import numpy as np

rng = np.random.default_rng(0)


def canon(r, b, kbytes):
    return (r // 8) * (kbytes * 8) + (b // 16) * 128 + (r % 8) * 16 + b % 16


def bake(mat_bytes, rows_alloc):
    """row-major [rows][kbytes] uint8 -> canonical image of rows_alloc rows (rows beyond the matrix = 0xEE garbage)"""
    rows, kb = mat_bytes.shape
    img = np.full(rows_alloc * kb, 0xEE, np.uint8)
    r, b = np.meshgrid(np.arange(rows), np.arange(kb), indexing="ij")
    img[canon(r, b, kb)] = mat_bytes
    return img


def read_tile(img, start, lbo, sbo, rows, kbytes=32):
    r, b = np.meshgrid(np.arange(rows), np.arange(kbytes), indexing="ij")
    return img[start + (r // 8) * sbo + (b // 16) * lbo + (r % 8) * 16 + b % 16]


def mma_i8(D, col, img_a, a_desc, img_b, b_desc, n, accumulate):
    a = read_tile(img_a, *a_desc, rows=128).view(np.int8).astype(np.int64)
    b = read_tile(img_b, *b_desc, rows=n).view(np.int8).astype(np.int64)
    prod = a @ b.T
    D[:, col:col + n] = (D[:, col:col + n] if accumulate else 0) + prod


def check_gru_layer(units=64, k_in=224, ts=8):
    rows = 3 * units; rows_alloc = 2 * units + 128; k_rec = units
    Wi = rng.integers(-127, 128, (rows, k_in), dtype=np.int8); Wr = rng.integers(-127, 128, (rows, k_rec), dtype=np.int8)
    x = rng.integers(-127, 128, (ts, k_in), dtype=np.int8); h = rng.integers(-127, 128, (ts, k_rec), dtype=np.int8)
    sWi, sWr = bake(Wi.view(np.uint8), rows_alloc), bake(Wr.view(np.uint8), rows_alloc)
    sX, sH = bake(x.view(np.uint8), ts), bake(h.view(np.uint8), ts)
    D = np.zeros((128, 64), np.int64)
    for g in range(3):                                      # umma_gru_layer.cu, issue loop
        ai = (g * units // 8) * (k_in * 8); ar = (g * units // 8) * (k_rec * 8)
        for k in range(k_in // 32):
            mma_i8(D, g * ts, sWi, (ai + k * 256, 128, k_in * 8), sX, (k * 256, 128, k_in * 8), ts, k != 0)
        for k in range(k_rec // 32):
            mma_i8(D, (3 + g) * ts, sWr, (ar + k * 256, 128, k_rec * 8), sH, (k * 256, 128, k_rec * 8), ts, k != 0)
    gi = Wi.astype(np.int64) @ x.astype(np.int64).T; gr = Wr.astype(np.int64) @ h.astype(np.int64).T       # [rows][ts]
    for g in range(3):
        assert np.array_equal(D[:units, g * ts:(g + 1) * ts], gi[g * units:(g + 1) * units]), ("gru input gate", g)
        assert np.array_equal(D[:units, (3 + g) * ts:(4 + g) * ts], gr[g * units:(g + 1) * units]), ("gru recurrent gate", g)
    print(f"gru layer  units={units} k_in={k_in}: gates land in lanes 0..{units - 1} of their column blocks")


def check_conv_layer(out=96, kc=288, ts=8):
    K = 2 * kc
    W = rng.integers(-127, 128, (out, K), dtype=np.int8)
    xo = rng.integers(-127, 128, (ts, kc), dtype=np.int8); xc = rng.integers(-127, 128, (ts, kc), dtype=np.int8)
    sW, sOld, sCur = bake(W.view(np.uint8), 128), bake(xo.view(np.uint8), ts), bake(xc.view(np.uint8), ts)
    D = np.zeros((128, 8), np.int64)
    for k in range(K // 32):                                # umma_conv_layer.cu, issue loop
        old = k < kc // 32
        b_img, b_start = (sOld, k * 256) if old else (sCur, (k - kc // 32) * 256)
        mma_i8(D, 0, sW, (k * 256, 128, K * 8), b_img, (b_start, 128, kc * 8), ts, k != 0)
    ref = W[:, :kc].astype(np.int64) @ xo.astype(np.int64).T + W[:, kc:].astype(np.int64) @ xc.astype(np.int64).T
    assert np.array_equal(D[:out], ref)
    print(f"conv layer out={out} taps=2x{kc}: old / current frame buffers accumulate into one chain")


def check_gru_chain(units=64, L=3, ts=8):
    rows = 3 * units; rows_alloc = 2 * units + 128; kcat = units * (L + 1)
    k_in = lambda l: units * (l + 1)
    wi_off = lambda l: rows_alloc * units * (l * (l + 1) // 2)
    Wi = [rng.integers(-127, 128, (rows, k_in(l)), dtype=np.int8) for l in range(L)]
    sWi = np.concatenate([bake(Wi[l].view(np.uint8), rows_alloc) for l in range(L)])
    assert all(wi_off(l) == sum(rows_alloc * k_in(j) for j in range(l)) for l in range(L + 1)), "wi_off"
    cat = rng.integers(-127, 128, (ts, kcat), dtype=np.int8)            # pretend every layer's output is already there
    sC = bake(cat.view(np.uint8), ts)
    cat_off = lambda s, k: (k // 16) * 128 + s * 16 + k % 16
    assert all(sC[cat_off(s, k)] == cat.view(np.uint8)[s, k] for s in range(ts) for k in range(0, kcat, 7)), "cat_off"
    for l in range(L):
        K = k_in(l); nkb = K // 32; fresh = 0 if l == 0 else 2
        D = np.zeros((128, 48), np.int64)
        for g in range(3):
            ai = wi_off(l) + (g * units // 8) * (K * 8)
            order = list(range(nkb - fresh)) + list(range(nkb - fresh, nkb))
            for k in order:
                mma_i8(D, g * ts, sWi, (ai + k * 256, 128, K * 8), sC, (k * 256, 128, kcat * 8), ts, k != 0)
        ref = Wi[l].astype(np.int64) @ cat[:, :K].astype(np.int64).T
        for g in range(3):
            assert np.array_equal(D[:units, g * ts:(g + 1) * ts], ref[g * units:(g + 1) * units]), ("chain", l, g)
    print(f"gru chain  L={L}: every layer reads its concat prefix through the shared B buffer")


def check_tf32_refresh(ntap=160, nf=40, nrow=48, kc=64):
    M, N, K = 128, 2 * nrow, 2 * ntap
    kcb = kc * 4
    A = np.zeros((M, K), np.float32); A[:2 * nf] = rng.standard_normal((2 * nf, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    D = np.zeros((M, N))
    for c in range(K // kc):                               # umma_tf32_refresh.cu: chunk staging + 8 k-steps per chunk
        sA = bake(np.ascontiguousarray(A[:, c * kc:(c + 1) * kc]).view(np.uint8), M)
        sB = bake(np.ascontiguousarray(B[:, c * kc:(c + 1) * kc]).view(np.uint8), N)
        for ks in range(kc // 8):
            a = read_tile(sA, ks * 256, 128, kcb * 8, rows=M).copy().view(np.float32).astype(np.float64)
            b = read_tile(sB, ks * 256, 128, kcb * 8, rows=N).copy().view(np.float32).astype(np.float64)
            D += a @ b.T
    assert np.allclose(D, A.astype(np.float64) @ B.astype(np.float64).T, rtol=0, atol=1e-9)
    print(f"tf32 refresh M={M} N={N} K={K} in {K // kc} chunks: chunk / k-step offsets consistent")


if __name__ == "__main__":
    check_gru_layer(); check_gru_layer(units=96, k_in=352)     # decoder shape: tiles start at rows 0, 96, 192
    check_conv_layer(); check_conv_layer(out=32, kc=192)
    check_gru_chain()
    check_tf32_refresh()
    print("all address checks passed")
