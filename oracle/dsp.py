"""NumPy restatement of the RADE V1 OFDM modem DSP chain — TEST INFRASTRUCTURE, NOT THE PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this.
Each function cites the reference code it restates (paths relative to /root/reference).  Pinned against
the reference itself by tools/make_golden.py (which imports /root/reference/radae*, runs both on the same
seeded inputs and commits the reference's outputs under tests/golden/).

Precision notes (what the reference computes in, which this file mirrors):
  * transmitter / receiver_one: torch complex64.
  * complex_bpf, detect_pilots, check_pilots row refresh: numpy csingle.
  * refine, check_pilots spot correlations, rx phase recursion: numpy complex128, results stored as csingle.
  * fmax tracking: Python float (float64).
One deliberate, documented deviation: check_pilots' 48 *random* row refreshes (radae/dsp.py:291-295, unseeded
np.random.randint) are replaced by the deterministic schedule `refresh_rows(n)`; only the per-row sums of |Dt| are
kept (they are all the reference ever reads back from the grids).
"""
import math
import numpy as np

# ---------------------------------------------------------------------------------------------
# constants: RADAE.__init__ (radae/radae.py:128-232) for model19_check3:
#   RADAE(21, 80, pilots=True, cyclic_prefix=0.004, bottleneck=3, time_offset=-16, coarse_mag=True)
# ---------------------------------------------------------------------------------------------
FS = 8000
M = 160
NCP = 32
NC = 30
NS = 4
NZMF = 3
LATENT = 80
NMF = (NS + 1) * (M + NCP)          # 960
NEOO = (NS + 2) * (M + NCP)         # 1152
NSEOO = (NS - 1) * NC               # 90
N_EOO_BITS = 2 * NSEOO              # 180
TIME_OFFSET = -16
RXBUF = 2 * NMF + M + NCP           # 2112
NIN_MAX = NMF + M                   # 1120
NB_TOTAL_FEATURES = 36
NUM_USED_FEATURES = 20
N_FEATURES = NZMF * 4 * NB_TOTAL_FEATURES   # 432
CARRIER_1_INDEX = 15
NMF_UNSYNC = int(3.0 * FS / NMF)    # 25   (radae_rxe.py:52,126)
SYNCED_ONE_SEC = FS // NMF          # 8    (radae_rxe.py:137)
UW_ERROR_THRESH = 7                 # radae_rxe.py:53
NUPDATE = int(0.05 * NMF)           # 48   (radae/dsp.py:290)
BARKER13 = np.array([1, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1, 1], np.float64)


class Consts:
    """Everything RADAE.__init__ derives (radae/radae.py:172-219), complex64 like the torch tensors."""

    def __init__(self):
        f32 = np.float32
        # self.w is a float32 torch tensor: 2*pi*(15+arange(Nc))/M  (radae.py:174)
        self.w = (f32(2 * math.pi) * (CARRIER_1_INDEX + np.arange(NC, dtype=f32)) / f32(M)).astype(f32)
        n = np.arange(M, dtype=f32)
        # Winv[c,:] = exp(1j*arange(M)*w[c])/M ; Wfwd[:,c] = exp(-1j*arange(M)*w[c])  (float32 products, radae.py:178-179)
        ang = (n[None, :] * self.w[:, None]).astype(f32)
        self.Winv = (np.exp(1j * ang.astype(np.float64)).astype(np.complex64) / f32(M)).astype(np.complex64)
        self.Wfwd = np.exp(-1j * ang.astype(np.float64)).astype(np.complex64).T.copy()
        self.P = (f32(2 ** 0.5) * BARKER13[np.arange(NC) % 13]).astype(np.complex64)            # radae.py:182
        self.Pend = self.P.copy(); self.Pend[1::2] *= -1                                       # radae.py:184-185
        self.p = (self.P @ self.Winv).astype(np.complex64)                                     # radae.py:183
        self.pend = (self.Pend @ self.Winv).astype(np.complex64)
        self.p_cp = np.concatenate([self.p[-NCP:], self.p])
        self.pend_cp = np.concatenate([self.pend[-NCP:], self.pend])
        self.pilot_gain = 10 ** (-2 / 20) * M / (NC ** 0.5)                                    # radae.py:195-199
        # acquisition tables (radae/dsp.py:153-176)
        self.fcoarse = np.arange(-50.0, 50.0, 2.5)
        self.p_w = np.zeros((M, len(self.fcoarse)), np.complex64)
        for i, f in enumerate(self.fcoarse):
            self.p_w[:, i] = np.exp(1j * (2 * np.pi * f / FS) * np.arange(M)) * self.p
        # LS pilot-EQ projectors (radae/dsp.py:401-412): plain transpose, NOT conjugate transpose
        a = 0.0025 * FS
        self.eq_a = a
        self.Pmat = np.zeros((NC, 2, 3), np.complex64)
        for c in range(NC):
            cm = min(max(c, 1), NC - 2)
            A = np.array([[1, np.exp(-1j * np.float64(self.w[cm - 1]) * a)],
                          [1, np.exp(-1j * np.float64(self.w[cm]) * a)],
                          [1, np.exp(-1j * np.float64(self.w[cm + 1]) * a)]]).astype(np.complex64)
            self.Pmat[c] = (np.linalg.inv((A.T @ A).astype(np.complex64)) @ A.T).astype(np.complex64)
        # torch.exp(-1j*self.w[c]*a) on a complex64 scalar: the angle w_c*a is a float32 product (dsp.py:433)
        self.eq_rot = np.exp(-1j * (self.w * f32(a)).astype(f32).astype(np.float64)).astype(np.complex64)
        # BPF (radae_rxe.py:104-109): `w = np.array(model.w)` is float32, so bandwidth and centre are evaluated in
        # float32 scalar arithmetic (NEP 50: Python literals adopt float32), left to right
        f32 = np.float32
        w0, w29 = f32(self.w[0]), f32(self.w[NC - 1])
        self.bpf_bw = f32(f32(f32(f32(1.2) * f32(w29 - w0)) * f32(FS)) / f32(2 * np.pi))
        self.bpf_centre = f32(f32(f32(f32(w29 + w0) * f32(FS)) / f32(2 * np.pi)) / f32(2))
        # EOO frame without data symbols: P E 0 0 0 E (radae.py:208-219)
        eoo = np.zeros(NEOO, np.complex64)
        eoo[:M + NCP] = self.p_cp
        eoo[M + NCP:2 * (M + NCP)] = self.pend_cp
        eoo[NMF:NMF + M + NCP] = self.pend_cp
        eoo = (eoo * np.float32(self.pilot_gain)).astype(np.complex64)
        self.eoo_base = pa_limiter(eoo)


def pa_limiter(x):
    """tanh(|x|)*exp(j*angle(x))  (radae/dsp.py:376-377), complex64"""
    x = x.astype(np.complex64)
    mag = np.abs(x).astype(np.float32)
    ang = np.angle(x).astype(np.float32)
    return (np.tanh(mag).astype(np.float32) * np.exp(1j * ang.astype(np.float64))).astype(np.complex64)


_C = None


def consts():
    global _C
    if _C is None:
        _C = Consts()
    return _C


# ---------------------------------------------------------------------------------------------
# transmitter (radae/dsp.py:340-378) and EOO (radae/radae.py:441-455)
# ---------------------------------------------------------------------------------------------
def transmitter_one(z):
    """z [3,80] float32 -> 960 complex64 samples of one modem frame."""
    c = consts()
    z = np.asarray(z, np.float32).reshape(NZMF, LATENT)
    sym = (z[:, 0::2] + 1j * z[:, 1::2]).astype(np.complex64).reshape(NS, NC)     # symbol k -> row k//30, carrier k%30
    frame = np.zeros((NS + 1, NC), np.complex64)
    frame[0] = (np.float32(c.pilot_gain) * c.P).astype(np.complex64)
    frame[1:] = sym
    tx = (frame @ c.Winv).astype(np.complex64)                                     # [5,160]
    tx = np.concatenate([tx[:, -NCP:], tx], axis=1).reshape(-1)                    # cyclic prefix = tail copy
    return pa_limiter(tx)


def eoo_frame(eoo_bits=None):
    """1152-sample end-of-over frame; eoo_bits = 180 floats (+-1) or None (zeros in the data slots)."""
    c = consts()
    out = c.eoo_base.copy()
    if eoo_bits is not None:
        b = np.asarray(eoo_bits, np.float32)
        syms = (b[0::2] + 1j * b[1::2]).astype(np.complex64).reshape(NS - 1, NC)
        tx = (syms @ c.Winv).astype(np.complex64)
        tx = np.concatenate([tx[:, -NCP:], tx], axis=1).reshape(-1)
        tx = (tx * np.float32(c.pilot_gain)).astype(np.complex64)
        out[2 * (M + NCP):NMF] = pa_limiter(tx)
    return out


# ---------------------------------------------------------------------------------------------
# complex band-pass filter (radae/dsp.py:39-102) including the 102-sample memory quirk (dsp.py:96)
# ---------------------------------------------------------------------------------------------
class ComplexBPF:
    NTAP = 101

    def __init__(self, bandwidth_hz=None, centre_hz=None, max_len=FS):
        c = consts()
        bw = c.bpf_bw if bandwidth_hz is None else bandwidth_hz
        fc = c.bpf_centre if centre_hz is None else centre_hz
        # complex_bpf.__init__ (radae/dsp.py:40-61) with float32 arguments: B, alpha and the taps are float32
        f32 = np.float32
        B = f32(f32(bw) / f32(FS))
        self.alpha = f32(f32(f32(2 * np.pi) * f32(fc)) / f32(FS))
        n = (np.arange(self.NTAP) - (self.NTAP - 1) / 2).astype(f32)
        y = (f32(np.pi) * np.where(n * B == 0, f32(1.0e-20), (n * B).astype(f32))).astype(f32)   # np.sinc in float32
        self.h = (B * (np.sin(y).astype(f32) / y).astype(f32)).astype(f32)    # real taps (stored csingle in the reference)
        # np.exp(-1j*alpha*arange(1,max_len+1), dtype=csingle): product in double, ARGUMENT rounded to float32, then exp
        arg = (-(np.float64(self.alpha) * np.arange(1, max_len + 1))).astype(f32)
        self.phase_vec_exp = np.exp(1j * arg.astype(np.float64)).astype(np.complex64)
        self.phase = np.complex64(1)
        self.mem = np.zeros(self.NTAP - 1, np.complex64)              # 100 on the first call, 102 afterwards

    def bpf(self, x):
        x = np.asarray(x, np.complex64)
        n = len(x)
        phase_vec = (self.phase * self.phase_vec_exp[:n]).astype(np.complex64)
        xb = (x * phase_vec).astype(np.complex64)
        x_mem = np.concatenate([self.mem, xb])
        win = np.lib.stride_tricks.sliding_window_view(x_mem, self.NTAP)[:n]
        y = (win @ self.h.astype(np.complex64)).astype(np.complex64)
        self.mem = x_mem[-self.NTAP - 1:].copy()
        self.phase = phase_vec[-1]
        return (y * np.conj(phase_vec)).astype(np.complex64)


# ---------------------------------------------------------------------------------------------
# acquisition (radae/dsp.py:152-320)
# ---------------------------------------------------------------------------------------------
def refresh_rows(n_call):
    """Deterministic replacement for the reference's 48 random rows per check_pilots call: 48 rows spaced 20
    samples apart, the offset rotating with the call counter, so all 960 rows refresh every 20 frames."""
    return (20 * np.arange(NUPDATE) + (n_call % 20)) % NMF


def arange_like_numpy(start, stop, step):
    """np.arange for float64 arguments, spelled out so the CUDA host code can restate it exactly:
    len = ceil((stop-start)/step); v[i] = start + i*((start+step)-start)."""
    n = int(math.ceil((stop - start) / step))
    delta = (start + step) - start
    return np.array([start + i * delta for i in range(n)], np.float64) if n > 0 else np.zeros(0)


class Acquisition:
    def __init__(self, Pacq_error1=1e-5, Pacq_error2=1e-4):
        self.c = consts()
        self.Pacq_error1, self.Pacq_error2 = Pacq_error1, Pacq_error2
        self.rowsum1 = np.zeros(NMF, np.float32)      # sum_f |Dt1[t,f]|
        self.rowsum2 = np.zeros(NMF, np.float32)
        self.Dthresh = 0.0; self.Dtmax12 = 0.0; self.Dtmax12_eoo = 0.0
        self.n_check = 0

    def _rows(self, rx_conj, ts):
        """Dt1[t,:], Dt2[t,:] for the given rows (complex64 matmul like np.matmul on csingle)."""
        idx = np.asarray(ts)[:, None] + np.arange(M)[None, :]
        D1 = (rx_conj[idx] @ self.c.p_w).astype(np.complex64)
        D2 = (rx_conj[idx + NMF] @ self.c.p_w).astype(np.complex64)
        return D1, D2

    def _sigma_r(self):
        s1 = np.float32(self.rowsum1.sum(dtype=np.float32) / np.float32(NMF * 40)) / np.float32((np.pi / 2) ** 0.5)
        s2 = np.float32(self.rowsum2.sum(dtype=np.float32) / np.float32(NMF * 40)) / np.float32((np.pi / 2) ** 0.5)
        return np.float32((s1 + s2) / np.float32(2.0))

    def detect_pilots(self, rx):
        """radae/dsp.py:178-231 -> (candidate, tmax, fmax)"""
        assert len(rx) == RXBUF
        rx_conj = np.conj(np.asarray(rx, np.complex64))
        D1, D2 = self._rows(rx_conj, np.arange(NMF))
        A1, A2 = np.abs(D1).astype(np.float32), np.abs(D2).astype(np.float32)
        D12 = (A1 + A2).astype(np.float32)
        # first strict maximum in (t, then f) order == the reference's running `local_max > Dtmax12`
        row_max = D12.max(axis=1)
        tmax = int(np.argmax(row_max))
        f_ind = int(np.argmax(D12[tmax]))
        Dtmax12 = float(row_max[tmax])
        if Dtmax12 <= 0:
            tmax, f_ind = 0, 0
        fmax = float(self.c.fcoarse[f_ind]) if Dtmax12 > 0 else 0.0
        self.rowsum1 = A1.sum(axis=1, dtype=np.float32)
        self.rowsum2 = A2.sum(axis=1, dtype=np.float32)
        sigma_r = self._sigma_r()
        self.Dthresh = float(2 * sigma_r * np.sqrt(-np.log(self.Pacq_error1 / 5.0)))
        self.Dtmax12 = Dtmax12
        return Dtmax12 > self.Dthresh, tmax, fmax

    def refine(self, rx, tmax, fmax, tfine_range, ffine_range):
        """radae/dsp.py:233-270: complex128 correlations stored as csingle; f outer loop, t inner, strict >."""
        p = self.c.p.astype(np.complex128)
        rx = np.asarray(rx, np.complex64)
        Dtmax = 0.0
        n = np.arange(M)
        for f in ffine_range:
            w = 2 * np.pi * f / FS
            w1 = np.exp(-1j * w * n) * np.conj(p)
            w2 = np.exp(-1j * w * n) * np.exp(-1j * w * NMF) * np.conj(p)
            for t in tfine_range:
                d1 = np.complex64(np.dot(rx[t:t + M].astype(np.complex128), w1))
                d2 = np.complex64(np.dot(rx[t + NMF:t + NMF + M].astype(np.complex128), w2))
                mag = np.abs(np.complex64(d1 + d2))
                if mag > Dtmax:
                    Dtmax = mag; tmax = int(t); fmax = float(f)
        return tmax, fmax

    def check_pilots(self, rx, tmax, fmax):
        """radae/dsp.py:273-320 with the deterministic row schedule -> (valid, endofover)"""
        c = self.c
        rx = np.asarray(rx, np.complex64)
        rx_conj = np.conj(rx)
        ts = refresh_rows(self.n_check); self.n_check += 1
        D1, D2 = self._rows(rx_conj, ts)
        self.rowsum1[ts] = np.abs(D1).astype(np.float32).sum(axis=1, dtype=np.float32)
        self.rowsum2[ts] = np.abs(D2).astype(np.float32).sum(axis=1, dtype=np.float32)
        sigma_r = self._sigma_r()
        Dthresh = float(2 * sigma_r * np.sqrt(-np.log(self.Pacq_error2 / 5.0)))
        Dthresh_eoo = float(2 * sigma_r * np.sqrt(-np.log(self.Pacq_error1 / 5.0)))
        w = 2 * np.pi * fmax / FS
        w_vec = np.exp(-1j * w * np.arange(M))
        p, pend = c.p.astype(np.complex128), c.pend.astype(np.complex128)
        r = rx.astype(np.complex128)
        D = abs(np.dot(np.conj(w_vec * r[tmax:tmax + M]), p)) + abs(np.dot(np.conj(w_vec * r[tmax + NMF:tmax + NMF + M]), p))
        De = abs(np.dot(np.conj(w_vec * r[tmax + M + NCP:tmax + 2 * M + NCP]), pend)) \
            + abs(np.dot(np.conj(w_vec * r[tmax + NMF:tmax + NMF + M]), pend))
        self.Dthresh, self.Dtmax12, self.Dtmax12_eoo = Dthresh, float(D), float(De)
        return D > Dthresh, De > Dthresh_eoo


# ---------------------------------------------------------------------------------------------
# receiver_one (radae/dsp.py:383-526)
# ---------------------------------------------------------------------------------------------
class ReceiverOne:
    SNR_M, SNR_C = 0.8070, 2.513

    def __init__(self):
        self.c = consts()
        self.snrdB_3k_est = 0.0

    def est_pilots(self, sym):
        """3-pilot LS fit per carrier on pilot rows 0 and 5 (dsp.py:418-435) -> [2,30] complex64"""
        c = self.c
        out = np.zeros((2, NC), np.complex64)
        for i, row in enumerate((0, NS + 1)):
            hp = (sym[row] / c.P).astype(np.complex64)
            for k in range(NC):
                cm = min(max(k, 1), NC - 2)
                g = (c.Pmat[k] @ hp[cm - 1:cm + 2]).astype(np.complex64)
                out[i, k] = np.complex64(g[0] + g[1] * c.eq_rot[k])
        return out

    def update_snr_est(self, sym, rx_pilots):
        """dsp.py:438-456"""
        Pcn = sym[0]
        ph = np.angle(rx_pilots[0]).astype(np.float32)
        Rcn = (Pcn * np.exp(-1j * ph.astype(np.float64)).astype(np.complex64)).astype(np.complex64)
        S1 = float(np.sum(np.abs(Pcn).astype(np.float32) ** 2, dtype=np.float32))
        S2 = float(np.sum(np.abs(Rcn.imag).astype(np.float32) ** 2, dtype=np.float32)) + 1e-12
        snr = S1 / (2 * S2) - 1
        if snr <= 0:
            snr = 0.1
        snrdB = (10 * np.log10(snr) - self.SNR_C) / self.SNR_M
        Rs = FS / M
        snrdB_3k = snrdB + 10 * math.log10(Rs * NC / 3000) + 10 * math.log10((M + NCP) / M)
        self.snrdB_3k_est = 0.9 * self.snrdB_3k_est + 0.1 * snrdB_3k

    def receiver_one(self, rx, endofover):
        """rx: 1152 complex64 (freq-corrected) -> z_hat [3,80] float32, or 180 EOO soft bits"""
        c = self.c
        rx = np.asarray(rx, np.complex64).reshape(NS + 2, M + NCP)
        rx_dash = rx[:, NCP + TIME_OFFSET:NCP + TIME_OFFSET + M]
        sym = (rx_dash @ c.Wfwd).astype(np.complex64)                 # [6,30]
        if not endofover:
            rx_pilots = self.est_pilots(sym)
            self.update_snr_est(sym, rx_pilots)
            k = np.arange(NS + 2, dtype=np.float32)[:, None]
            slope = ((rx_pilots[1] - rx_pilots[0]) / np.float32(NS + 1)).astype(np.complex64)
            rx_ch = (slope[None, :] * k + rx_pilots[0][None, :]).astype(np.complex64)
            ang = np.angle(rx_ch).astype(np.float32)
            sym = sym.copy()
            sym[1:NS + 1] = (sym[1:NS + 1] * np.exp(-1j * ang[1:NS + 1].astype(np.float64)).astype(np.complex64)).astype(np.complex64)
            mag = np.float32(np.mean(np.abs(rx_pilots).astype(np.float32) ** 2, dtype=np.float32) ** np.float32(0.5)) + np.float32(1e-6)
            mag = np.float32(mag * np.abs(c.P[0]) / np.float32(c.pilot_gain))
            sym = (sym / mag).astype(np.complex64)
            d = sym[1:NS + 1].reshape(NZMF, LATENT // 2)
            z = np.zeros((NZMF, LATENT), np.float32)
            z[:, 0::2] = d.real; z[:, 1::2] = d.imag
            return z
        # end of over: average of the three pilot-ish symbols P, E, E (dsp.py:513-524)
        acc = (sym[0] / c.P + sym[1] / c.Pend + sym[NS + 1] / c.Pend).astype(np.complex64)
        ph = np.angle(acc).astype(np.float32)
        sym = (sym * np.exp(-1j * ph.astype(np.float64)).astype(np.complex64)[None, :]).astype(np.complex64)
        d = sym[2:NS + 1].reshape(-1)
        z = np.zeros(2 * d.size, np.float32)
        z[0::2] = d.real; z[1::2] = d.imag
        return z


# ---------------------------------------------------------------------------------------------
# streaming TX / RX objects (radae_txe.py:47-144, radae_rxe.py:56-330) on the int8 C core codec
# ---------------------------------------------------------------------------------------------
class RadaeTx:
    """One stream. `core` = a CoreOraclePort/CoreOracleRef with n_streams == 1 (the C encoder path,
    src/rade_api.c:411-434)."""

    def __init__(self, core, txbpf_en=False):
        self.core = core
        self.eoo_bits = None
        self.txbpf = ComplexBPF() if txbpf_en else None      # same band as the receive filter (radae_txe.py:74-81)

    def _bpf_clip(self, tx):
        """optional TX filter + unit-magnitude clip (radae_txe.py:130-132, :141-143), all in float32 / complex64"""
        if self.txbpf is None:
            return tx
        y = self.txbpf.bpf(tx)
        mag = np.abs(y).astype(np.float32)
        ang = np.angle(y).astype(np.float32)
        rot = (np.cos(ang).astype(np.float32) + 1j * np.sin(ang).astype(np.float32)).astype(np.complex64)
        return (np.clip(mag, 0, 1).astype(np.float32) * rot).astype(np.complex64)

    def do_radae_tx(self, features432):
        f = np.asarray(features432, np.float32).reshape(1, 12, NB_TOTAL_FEATURES)
        x = np.concatenate([f[:, :, :NUM_USED_FEATURES], -np.ones((1, 12, 1), np.float32)], axis=2).reshape(1, 3, 84)
        z = self.core.encode(x)[0]
        return self._bpf_clip(transmitter_one(z)), z

    def do_radae_tx_from_z(self, z240):
        """bypass_enc=True: the caller ran the core encoder (src/rade_api.c:411-434 -> radae_txe.py:122-124)"""
        return self._bpf_clip(transmitter_one(np.asarray(z240, np.float32).reshape(NZMF, LATENT)))

    def set_eoo_bits(self, bits):
        self.eoo_bits = np.asarray(bits, np.float32).copy()

    def do_eoo(self):
        return self._bpf_clip(eoo_frame(self.eoo_bits))


SEARCH, CANDIDATE, SYNC = 0, 1, 2


class RadaeRx:
    """One stream; restates radae_rx.do_radae_rx step by step.
    reset_dec_on_sync mirrors `model.core_decoder_statefull.module.reset()` (radae_rxe.py:263); the C-decoder
    path (RADE_USE_C_DECODER) keeps its own state and is never reset (src/rade_api.c:494-506)."""

    def __init__(self, core, foff_err=0.0, reset_dec_on_sync=False, bpf_en=True):
        self.core = core
        self.c = consts()
        self.bpf = ComplexBPF() if bpf_en else None
        self.acq = Acquisition()
        self.receiver = ReceiverOne()
        self.foff_err = foff_err
        self.reset_dec_on_sync = reset_dec_on_sync
        self.nin = NMF
        self.state = SEARCH
        self.tmax = 0; self.fmax = 0.0
        self.tmax_candidate = 0
        self.valid_count = 0
        self.uw_errors = 0
        self.synced_count = 0
        self.rx_phase = 1 + 0j
        self.rx_buf = np.zeros(RXBUF, np.complex64)
        self.z_hat = None

    def get_sync(self):
        return self.state == SYNC

    def get_snrdB_3k_est(self):
        return int(self.receiver.snrdB_3k_est)

    def do_radae_rx(self, samples):
        """samples: self.nin complex64 -> (ret, features432 or None, eoo_bits180 or None)"""
        acq = self.acq
        x = np.asarray(samples, np.complex64)[:self.nin]
        assert len(x) == self.nin
        valid_output = False; endofover = False; uw_fail = False
        if self.bpf is not None:
            x = self.bpf.bpf(x)
        self.rx_buf[:-self.nin] = self.rx_buf[self.nin:]
        self.rx_buf[-self.nin:] = x
        z_hat = None
        if self.state in (SEARCH, CANDIDATE):
            candidate, self.tmax, self.fmax = acq.detect_pilots(self.rx_buf)
        else:
            ffine = arange_like_numpy(self.fmax - 1, self.fmax + 1, 0.1)
            tfine = np.arange(max(0, self.tmax - 8), self.tmax + 8)
            self.tmax, fmax_hat = acq.refine(self.rx_buf, self.tmax, self.fmax, tfine, ffine)
            self.fmax = 0.9 * self.fmax + 0.1 * fmax_hat
            candidate, endofover = acq.check_pilots(self.rx_buf, self.tmax, self.fmax)
            self.nin = NMF
            if self.tmax >= NMF - M:
                self.nin = NMF + M; self.tmax -= M
            if self.tmax < M:
                self.nin = NMF - M; self.tmax += M
            self.synced_count += 1
            if self.synced_count % SYNCED_ONE_SEC == 0:
                if self.uw_errors > UW_ERROR_THRESH:
                    uw_fail = True
                self.uw_errors = 0
            w = 2 * np.pi * self.fmax / FS
            step = np.exp(-1j * w)
            vec = np.zeros(NEOO, np.complex64)
            ph = self.rx_phase
            for n in range(NEOO):
                ph = ph * step
                vec[n] = ph
            self.rx_phase = ph
            rx1 = self.rx_buf[self.tmax - NCP:self.tmax - NCP + NEOO]
            rx = (rx1 * vec).astype(np.complex64)            # csingle * csingle (radae_rxe.py:232-233)
            z_hat = self.receiver.receiver_one(rx, endofover)
            valid_output = not endofover

        prev = self.state
        nxt = self.state
        if self.state == SEARCH:
            if candidate:
                nxt = CANDIDATE; self.tmax_candidate = self.tmax; self.valid_count = 1
        elif self.state == CANDIDATE:
            if candidate and abs(self.tmax - self.tmax_candidate) < NCP:
                self.valid_count += 1
                if self.valid_count > 3:
                    nxt = SYNC
                    if self.reset_dec_on_sync:
                        self.core.dec_state[...] = 0
                    self.synced_count = 0; uw_fail = False; self.uw_errors = 0
                    self.valid_count = NMF_UNSYNC
                    ffine = arange_like_numpy(self.fmax - 10, self.fmax + 10, 0.25)
                    tfine = np.arange(max(0, self.tmax - 1), self.tmax + 2)
                    self.tmax, self.fmax = acq.refine(self.rx_buf, self.tmax, self.fmax, tfine, ffine)
                    self.fmax += self.foff_err; self.foff_err = 0
            else:
                nxt = SEARCH
        else:
            if candidate:
                self.valid_count = NMF_UNSYNC
            else:
                self.valid_count -= 1
                if self.valid_count == 0:
                    nxt = SEARCH
            if endofover or uw_fail:
                nxt = SEARCH
        self.state = nxt
        if self.state == SEARCH:
            self.nin = NMF

        features = None; eoo = None
        if valid_output:
            f = self.core.decode(z_hat.reshape(1, 3, 80))[0].reshape(12, 21)
            self.uw_errors += int(np.sum(f[0::4, 20] > 0))
            out = np.zeros((12, NB_TOTAL_FEATURES), np.float32)
            out[:, :20] = f[:, :20]
            features = out.reshape(-1)
        if endofover:
            eoo = z_hat.astype(np.float32)
        self.z_hat = z_hat
        self.prev_state = prev
        return int(valid_output) | (int(endofover) << 1), features, eoo


# ---------------------------------------------------------------------------------------------
# channel simulator: rate-Fs branch of RADAE.forward (radae/radae.py:529-599), made per-stream
# ---------------------------------------------------------------------------------------------
def ebno_sigma(EbNodB):
    """sigma = sqrt(Fs/(EbNo*Rb)), Rb = latent_dim/Tz = 2000 (radae.py:131, :570-574)"""
    return math.sqrt(FS / (10 ** (EbNodB / 10) * (LATENT / 0.04)))


def mp_gain_of(tx, G1, G2, d):
    """the reference's power normalisation through the multipath model (radae.py:536-539), for ONE stream (the reference takes
    the means over its whole batch tensor; the batched implementations here normalise per stream)"""
    tx = np.asarray(tx, np.complex64)
    mp = (tx * G1).astype(np.complex64)
    if d:
        mp[d:] += (tx[:-d] * G2[:-d]).astype(np.complex64)
    else:
        mp += (tx * G2).astype(np.complex64)
    return float(np.sqrt(np.mean(np.abs(tx) ** 2) / np.mean(np.abs(mp) ** 2)))


def channel(tx, G1, G2, d, mp_gain, freq_offset, phase0, sigma, noise, gain=1.0, df_dt=0.0):
    """tx [T] c64; G1,G2 [T] c64 path gains; delay d samples; noise [T] c64 unit-variance complex normal.
    rx = gain*( mp_gain*(tx*G1 + shift_d(tx*G2)) * exp(j*(phase0 + cumsum(omega))) + sigma*noise ),
    omega[n] = 2*pi*(freq_offset + df_dt*n/Fs)/Fs (radae.py:546-550); pinned by tests/golden/channel.npz (RADAE.forward itself)"""
    tx = np.asarray(tx, np.complex64)
    mp = (tx * G1).astype(np.complex64)
    mp[d:] += (tx[:-d] * G2[:-d]).astype(np.complex64) if d else (tx * G2).astype(np.complex64)
    n = np.arange(len(tx), dtype=np.float64)
    ph = phase0 + 2 * np.pi / FS * (freq_offset * (n + 1) + df_dt * n * (n + 1) / (2 * FS))
    lin = np.exp(1j * ph).astype(np.complex64)
    return (np.float32(gain) * (np.float32(mp_gain) * mp * lin + np.float32(sigma) * noise)).astype(np.complex64)
