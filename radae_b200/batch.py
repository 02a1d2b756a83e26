"""Host-side (Python) face of the batched C ABI: S independent RADE streams on one B200.

Thin by design — sizes, pointer plumbing and error checks only; all arithmetic happens in libradae_b200.so.
numpy arrays are passed as host pointers (the library copies H2D/D2H itself); torch CUDA tensors, when the
caller has them, go through the `_dev` methods as raw device pointers.
"""
import ctypes as C
import numpy as np
from . import _capi as capi
from ._capi import NMF, NEOO, NIN_MAX, NFEAT, NEOO_BITS


def _np(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, a.ctypes.data


def pinned_empty(shape, dtype):
    """numpy array in pinned host memory (rade_b200_host_alloc); the memory lives as long as the array (never freed: small, few)"""
    lib = capi.lib()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.rade_b200_host_alloc(max(n, 16))
    if not p:
        raise RuntimeError("rade_b200_host_alloc failed")
    buf = (C.c_char * max(n, 16)).from_address(p)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def _check(rc, what):
    if rc is None or rc < 0:
        raise RuntimeError(f"libradae_b200: {what} failed (see stderr for the CUDA error)")
    return rc


class RadeBatch:
    def __init__(self, n_streams, device=-1, flags=capi.RADE_USE_C_ENCODER | capi.RADE_USE_C_DECODER | capi.RADE_VERBOSE_0,
                 weights=None):
        self.lib = capi.lib()
        self.S = int(n_streams)
        buf, n = (None, 0)
        if weights is not None:
            self._w = bytes(weights); buf, n = C.c_char_p(self._w), len(self._w)
        self.h = self.lib.rade_b200_open(self.S, device, flags, buf, n)
        if not self.h:
            raise RuntimeError("rade_b200_open failed: no usable sm_100 CUDA device or bad weights (no CPU fallback)")

    def close(self):
        if getattr(self, "h", None):
            self.lib.rade_b200_close(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing
    @property
    def cuda_stream(self):
        return self.lib.rade_b200_cuda_stream(self.h)

    def synchronize(self):
        _check(self.lib.rade_b200_synchronize(self.h), "synchronize")

    def launch_count(self):
        return int(self.lib.rade_b200_launch_count(self.h))

    def reset(self):
        _check(self.lib.rade_b200_reset(self.h), "reset")

    # ---- core codec (host arrays)
    def core_dims(self):
        """(input_dim, output_dim) of the loaded core codec on the host side: (84, 84), or (80, 80) for a model without the aux symbol"""
        i, o = C.c_int(0), C.c_int(0)
        _check(self.lib.rade_b200_core_dims(self.h, C.byref(i), C.byref(o)), "core_dims")
        return i.value, o.value

    def core_encode(self, features):
        f, pf = _np(features, np.float32)
        S, T, K = f.shape
        assert S == self.S and K == self.core_dims()[0]
        z = np.empty((S, T, 80), np.float32)
        _check(self.lib.rade_b200_core_encode(self.h, z.ctypes.data, pf, T), "core_encode")
        return z

    def core_decode(self, z):
        zz, pz = _np(z, np.float32)
        S, T, K = zz.shape
        assert S == self.S and K == 80
        f = np.empty((S, T, self.core_dims()[1]), np.float32)
        _check(self.lib.rade_b200_core_decode(self.h, f.ctypes.data, pz, T), "core_decode")
        return f

    # ---- transmitter
    def tx(self, features_in, out=None):
        f, pf = _np(features_in, np.float32)
        assert f.size == self.S * NFEAT
        if out is None:
            out = np.empty((self.S, NMF), np.complex64)
        _check(self.lib.rade_b200_tx(self.h, out.ctypes.data, pf), "tx")
        return out

    def tx_z(self, z):
        """bypass_enc transmitter: z [S][3][80] from the caller's core encoder -> tx [S][960]"""
        zz, pz = _np(z, np.float32)
        assert zz.size == self.S * 240
        out = np.empty((self.S, NMF), np.complex64)
        _check(self.lib.rade_b200_tx_z(self.h, out.ctypes.data, pz), "tx_z")
        return out

    def tx_bpf_enable(self, on=True):
        """radae_tx(txbpf_en=True): band-pass filter + clip on every transmitted frame (restarts the filter)"""
        _check(self.lib.rade_b200_tx_bpf_enable(self.h, int(on)), "tx_bpf_enable")

    def tx_set_eoo_bits(self, bits):
        b, pb = _np(bits, np.float32)
        assert b.size == self.S * NEOO_BITS
        _check(self.lib.rade_b200_tx_set_eoo_bits(self.h, pb), "tx_set_eoo_bits")

    def tx_eoo(self):
        out = np.empty((self.S, NEOO), np.complex64)
        _check(self.lib.rade_b200_tx_eoo(self.h, out.ctypes.data), "tx_eoo")
        return out

    # ---- receiver
    def nin(self):
        n = np.empty(self.S, np.int32)
        _check(self.lib.rade_b200_nin(self.h, n.ctypes.data), "nin")
        return n

    def rx(self, rx_in, active=None):
        """rx_in [S][1120] complex64 (row s: nin[s] fresh samples) -> (features [S][432], ret [S], eoo [S][180])"""
        x, px = _np(rx_in, np.complex64)
        assert x.shape == (self.S, NIN_MAX)
        feats = np.zeros((self.S, NFEAT), np.float32)
        ret = np.zeros(self.S, np.int32)
        eoo = np.zeros((self.S, NEOO_BITS), np.float32)
        pa = None
        if active is not None:
            a, pa = _np(active, np.uint8)
        _check(self.lib.rade_b200_rx(self.h, feats.ctypes.data, ret.ctypes.data, eoo.ctypes.data, px, pa), "rx")
        return feats, ret, eoo

    def rx_status(self):
        arr = (capi.RxStatus * self.S)()
        _check(self.lib.rade_b200_rx_get_status(self.h, arr), "rx_get_status")
        return list(arr)

    def rx_z_hat(self):
        z = np.empty((self.S, 240), np.float32)
        _check(self.lib.rade_b200_rx_get_z_hat(self.h, z.ctypes.data), "rx_get_z_hat")
        return z

    # ---- raw device-pointer entry points (ints = CUDA device addresses)
    def core_encode_dev(self, d_z, d_features, n_steps):
        _check(self.lib.rade_b200_core_encode_dev(self.h, d_z, d_features, n_steps), "core_encode_dev")

    def core_decode_dev(self, d_features, d_z, n_steps):
        _check(self.lib.rade_b200_core_decode_dev(self.h, d_features, d_z, n_steps), "core_decode_dev")

    def tx_dev(self, d_tx_out, d_features_in):
        _check(self.lib.rade_b200_tx_dev(self.h, d_tx_out, d_features_in), "tx_dev")

    def tx_z_dev(self, d_tx_out, d_z):
        _check(self.lib.rade_b200_tx_z_dev(self.h, d_tx_out, d_z), "tx_z_dev")

    def ofdm_mod_dev(self, d_tx_out, d_z):
        _check(self.lib.rade_b200_ofdm_mod_dev(self.h, d_tx_out, d_z), "ofdm_mod_dev")

    def rx_dev(self, d_features_out, d_ret, d_eoo_out, d_rx_in, d_active=None):
        _check(self.lib.rade_b200_rx_dev(self.h, d_features_out, d_ret, d_eoo_out, d_rx_in, d_active), "rx_dev")

    def channel_config(self, EbNodB=100.0, freq_offset_hz=0.0, freq_offset_spread_hz=0.0, doppler_spread_hz=0.0,
                       delay_samples=16, gain=1.0, seed=1):
        cfg = capi.ChannelCfg(EbNodB, freq_offset_hz, freq_offset_spread_hz, doppler_spread_hz, delay_samples, gain, seed)
        _check(self.lib.rade_b200_channel_config(self.h, C.byref(cfg)), "channel_config")

    def channel_dev(self, d_rx, d_tx):
        _check(self.lib.rade_b200_channel_dev(self.h, d_rx, d_tx), "channel_dev")

    def channel_apply_dev(self, d_rx, d_tx, d_G1, d_G2, d_noise, n, delay, mp_gain, freq, phase0, sigma, gain):
        _check(self.lib.rade_b200_channel_apply_dev(self.h, d_rx, d_tx, d_G1, d_G2, d_noise, n, delay, mp_gain, freq,
                                                    phase0, sigma, gain), "channel_apply_dev")

    def loopback_run(self, features_in, n_frames, features_out, ret, valid_frames=None):
        """features -> encoder -> modulator -> channel -> receiver -> decoder -> features for n_frames modem frames, host arrays at
        both ends (features_in [n_in][S][432] cycled over; features_out [S][432], ret [S] int32 overwritten every frame)"""
        f, pf = _np(features_in, np.float32)
        assert f.ndim == 3 and f.shape[1:] == (self.S, NFEAT) and features_out.dtype == np.float32 and ret.dtype == np.int32
        vp = valid_frames.ctypes.data if valid_frames is not None else None
        _check(self.lib.rade_b200_loopback_run(self.h, pf, f.shape[0], int(n_frames), features_out.ctypes.data, ret.ctypes.data, vp), "loopback_run")

    def loopback_step_dev(self, d_features_next, d_features_out, d_ret, d_eoo_out):
        _check(self.lib.rade_b200_loopback_step_dev(self.h, d_features_next, d_features_out, d_ret, d_eoo_out), "loopback_step_dev")

    def pipeline_enable(self, on=True):
        _check(self.lib.rade_b200_pipeline_enable(self.h, int(on)), "pipeline_enable")

    def pipeline_fork(self):
        _check(self.lib.rade_b200_pipeline_fork(self.h), "pipeline_fork")

    def pipeline_join(self):
        _check(self.lib.rade_b200_pipeline_join(self.h), "pipeline_join")

    def channel_link_dev(self, d_tx):
        _check(self.lib.rade_b200_channel_link_dev(self.h, d_tx), "channel_link_dev")

    def tx_channel_link_dev(self, d_features_in):
        _check(self.lib.rade_b200_tx_channel_link_dev(self.h, d_features_in), "tx_channel_link_dev")

    def rx_link_dev(self, d_features_out, d_ret, d_eoo_out):
        _check(self.lib.rade_b200_rx_link_dev(self.h, d_features_out, d_ret, d_eoo_out), "rx_link_dev")

    def channel_apply(self, tx, G1, G2, noise, delay=16, mp_gain=1.0, freq_offset_hz=0.0, phase0=0.0, sigma=0.0, gain=1.0, df_dt=0.0):
        """explicit channel on host arrays [S][n] complex64 (RADAE.forward rate-Fs branch, radae/radae.py:529-599):
        G1, G2 e.g. from a fading file (radae_b200.gfile.read_g), noise unit-variance complex normal"""
        arrs = [np.ascontiguousarray(a, np.complex64) for a in (tx, G1, G2, noise)]
        S, n = arrs[0].shape
        assert S == self.S and all(a.shape == (S, n) for a in arrs)
        out = np.empty((S, n), np.complex64)
        _check(self.lib.rade_b200_channel_apply_drift(self.h, out.ctypes.data, *[a.ctypes.data for a in arrs], n, delay, mp_gain,
                                                      freq_offset_hz, df_dt, phase0, sigma, gain), "channel_apply")
        return out

    def link_push_dev(self, d_samples):
        _check(self.lib.rade_b200_link_push_dev(self.h, d_samples), "link_push_dev")

    def link_pop_dev(self, d_rx_in, d_active):
        _check(self.lib.rade_b200_link_pop_dev(self.h, d_rx_in, d_active), "link_pop_dev")

    def channel(self, tx, out=None):
        """host-buffer channel call: tx [S][960] complex64 -> rx [S][960]"""
        t, pt = _np(tx, np.complex64)
        assert t.shape == (self.S, NMF)
        if out is None:
            out = np.empty((self.S, NMF), np.complex64)
        _check(self.lib.rade_b200_channel(self.h, out.ctypes.data, pt), "channel")
        return out

    # ---- per-kernel device timing
    def profile_enable(self, on=True):
        _check(self.lib.rade_b200_profile_enable(self.h, int(on)), "profile_enable")

    def profile_read(self):
        n = self.lib.rade_b200_profile_n_kernels()
        ms = np.zeros(n, np.float32); cnt = np.zeros(n, np.int32)
        _check(self.lib.rade_b200_profile_read(self.h, ms.ctypes.data, cnt.ctypes.data), "profile_read")
        return {self.lib.rade_b200_profile_kernel_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n) if cnt[k]}


class HostLink:
    """Pinned per-stream sample FIFOs in front of the receiver (rade_b200_hostlink_*): push 960 samples per stream,
    rx() advances every stream that has nin[s] samples queued.  Output arrays are allocated once and reused."""

    def __init__(self, batch, capacity=4096):
        self.b, self.lib = batch, batch.lib
        self.h = self.lib.rade_b200_hostlink_open(batch.h, capacity)
        if not self.h:
            raise RuntimeError("rade_b200_hostlink_open failed")
        S = batch.S
        self.features = pinned_empty((S, NFEAT), np.float32)          # pinned: the D2H copies run on the copy engine
        self.ret = pinned_empty((S,), np.int32)
        self.eoo = pinned_empty((S, NEOO_BITS), np.float32)

    def push(self, samples):
        """returns the number of streams whose frame was dropped because their FIFO was full"""
        x, px = _np(samples, np.complex64)
        assert x.shape == (self.b.S, NMF)
        return _check(self.lib.rade_b200_hostlink_push(self.h, px), "hostlink_push")

    def channel_push(self, bch, tx):
        """tx [S][960] -> channel simulator of context `bch` -> the next frame slot (kernel writes in place); 0 or S (dropped)"""
        t, pt = _np(tx, np.complex64)
        assert t.shape == (self.b.S, NMF)
        return _check(self.lib.rade_b200_channel_hostlink(bch.h, self.h, pt), "channel_hostlink")

    def duplex_run(self, btx, bch, features_in, n_frames, tx_bufs, valid_frames=None):
        """the reference's radae_tx | ch | radae_rx pipe for S streams, driven from C (three host threads, synchronous calls)"""
        f, pf = _np(features_in, np.float32)
        assert f.ndim == 3 and f.shape[1:] == (self.b.S, NFEAT)
        assert tx_bufs.dtype == np.complex64 and tx_bufs.ndim == 3 and tx_bufs.shape[1:] == (self.b.S, NMF) and tx_bufs.flags.c_contiguous
        vp = valid_frames.ctypes.data if valid_frames is not None else None
        _check(self.lib.rade_b200_duplex_run(btx.h, bch.h, self.h, pf, f.shape[0], int(n_frames), tx_bufs.ctypes.data, tx_bufs.shape[0],
                                             self.features.ctypes.data, self.ret.ctypes.data, vp), "duplex_run")
        return self.features, self.ret

    def rx(self):
        _check(self.lib.rade_b200_hostlink_rx(self.h, self.features.ctypes.data, self.ret.ctypes.data, self.eoo.ctypes.data), "hostlink_rx")
        return self.features, self.ret, self.eoo

    def dropped(self):
        """frames refused by push() so far because a FIFO was full"""
        return int(self.lib.rade_b200_hostlink_dropped(self.h))

    def close(self):
        if self.h:
            self.lib.rade_b200_hostlink_close(self.h); self.h = None


class RadeMulti:
    """S streams over several GPUs from one process (rade_b200_open_multi / rade_b200_open_devices): contiguous blocks of
    streams per device, calls routed by block with one host thread per device; arrays as RadeBatch's with S = n_streams."""

    def __init__(self, n_streams, devices=None, device_mask=0, flags=capi.RADE_USE_C_ENCODER | capi.RADE_USE_C_DECODER | capi.RADE_VERBOSE_0,
                 weights=None):
        self.lib = capi.lib()
        self.S = int(n_streams)
        buf, n = (None, 0)
        if weights is not None:
            self._w = bytes(weights); buf, n = C.c_char_p(self._w), len(self._w)
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            self.h = self.lib.rade_b200_open_devices(self.S, arr, len(devices), flags, buf, n)
        else:
            self.h = self.lib.rade_b200_open_multi(self.S, device_mask, flags, buf, n)
        if not self.h:
            raise RuntimeError("rade_b200_open_multi failed: no usable sm_100 CUDA device, bad device list or bad weights")

    @property
    def n_devices(self):
        return int(self.lib.rade_b200_multi_n_devices(self.h))

    def blocks(self):
        out = []
        for i in range(self.n_devices):
            f, c = C.c_int(0), C.c_int(0)
            self.lib.rade_b200_multi_context(self.h, i, C.byref(f), C.byref(c))
            out.append((f.value, c.value))
        return out

    def tx(self, features_in):
        f, pf = _np(features_in, np.float32)
        assert f.size == self.S * NFEAT
        out = np.empty((self.S, NMF), np.complex64)
        _check(self.lib.rade_b200_multi_tx(self.h, out.ctypes.data, pf), "multi_tx")
        return out

    def nin(self):
        n = np.zeros(self.S, np.int32)
        _check(self.lib.rade_b200_multi_nin(self.h, n.ctypes.data), "multi_nin")
        return n

    def rx(self, rx_in, active=None):
        x, px = _np(rx_in, np.complex64)
        assert x.shape == (self.S, NIN_MAX)
        f = np.zeros((self.S, NFEAT), np.float32); ret = np.zeros(self.S, np.int32); eoo = np.zeros((self.S, NEOO_BITS), np.float32)
        pa = None
        if active is not None:
            a, pa = _np(active, np.uint8)
        _check(self.lib.rade_b200_multi_rx(self.h, f.ctypes.data, ret.ctypes.data, eoo.ctypes.data, px, pa), "multi_rx")
        return f, ret, eoo

    def close(self):
        if getattr(self, "h", None):
            self.lib.rade_b200_close_multi(self.h); self.h = None
