"""Host-side mirror of the reference's streaming objects, on the C ABI (and through it on the B200).

Same class names, constructor keywords, method names and return conventions as
  radae_txe.radae_tx  (/root/reference/radae_txe.py:47-144)
  radae_rxe.radae_rx  (/root/reference/radae_rxe.py:56-330)
so a script written against the reference's modules runs against these by changing the import.  Unsupported
reference options raise instead of silently doing something else.  `model_name` is a weight file (RDW or DNNw) or
anything else for the embedded model19_check3 weights — mirroring rade_open's model_file handling.
"""
import ctypes as C
import os
import numpy as np
from . import _capi as capi

nb_total_features = 36
num_used_features = 20


def _open(model_name, flags):
    lib = capi.lib()
    lib.rade_initialize()
    h = lib.rade_open((model_name or "embedded").encode(), flags)
    if not h:
        raise RuntimeError("rade_open failed")
    return lib, h


def _weights_of(model_name):
    """rade_open's model_file rule (csrc/api.cu): a readable RDW / DNNw file is used, anything else means the embedded weights"""
    if model_name and os.path.isfile(model_name):
        with open(model_name, "rb") as f:
            blob = f.read()
        if blob[:8] == b"RADEB200" or blob[:4] == b"DNNw":
            return blob
    return None


_FLAGS = capi.RADE_USE_C_ENCODER | capi.RADE_USE_C_DECODER | capi.RADE_VERBOSE_0


class radae_tx:
    """radae_txe.radae_tx.  Default options go through the reference's own single-stream surface (rade_open / rade_tx / ...);
    bypass_enc (the caller supplies the 3 x 80 latents of its own core encoder, radae_txe.py:122-124) and txbpf_en (TX
    band-pass filter + clip, radae_txe.py:74-81, :130-132) are not reachable through rade_api.h and use a 1-stream batch."""

    def __init__(self, model_name=None, latent_dim=80, auxdata=True, bottleneck=3, txbpf_en=False, bypass_enc=False):
        if latent_dim != 80 or not auxdata or bottleneck != 3:
            raise NotImplementedError("libradae_b200 implements the RADE V1 waveform: latent_dim=80, auxdata, bottleneck=3")
        self.bypass_enc, self.txbpf_en = bool(bypass_enc), bool(txbpf_en)
        self.h = self.batch = None
        if self.bypass_enc or self.txbpf_en:
            from .batch import RadeBatch
            self.batch = RadeBatch(1, flags=_FLAGS, weights=_weights_of(model_name))
            if self.txbpf_en:
                self.batch.tx_bpf_enable(True)
        else:
            self.lib, self.h = _open(model_name, _FLAGS)
        self.n_floats_in = capi.NZMF * capi.LATENT if self.bypass_enc else capi.NFEAT
        self.Nmf, self.Neoo = capi.NMF, capi.NEOO

    def get_n_features_in(self): return capi.NFEAT
    def get_n_floats_in(self): return self.n_floats_in
    def get_Nmf(self): return self.Nmf
    def get_Neoo(self): return self.Neoo
    def get_Neoo_bits(self): return capi.NEOO_BITS

    def set_eoo_bits(self, eoo_bits):
        b = np.ascontiguousarray(eoo_bits, np.float32)
        assert b.size == self.get_Neoo_bits()
        if self.batch:
            self.batch.tx_set_eoo_bits(b)
        else:
            self.lib.rade_tx_set_eoo_bits(self.h, b.ctypes.data)

    def do_radae_tx(self, buffer_f32, tx_out):
        f = np.ascontiguousarray(buffer_f32, np.float32)
        assert f.size == self.n_floats_in and tx_out.dtype == np.complex64 and tx_out.size == self.Nmf
        if self.batch:
            out = (self.batch.tx_z(f) if self.bypass_enc else self.batch.tx(f))[0]
        else:
            out = np.empty(self.Nmf, np.complex64)
            self.lib.rade_tx(self.h, out.ctypes.data, f.ctypes.data)
        np.copyto(tx_out, out)

    def do_eoo(self, tx_out):
        if self.batch:
            out = self.batch.tx_eoo()[0]
        else:
            out = np.empty(self.Neoo, np.complex64)
            self.lib.rade_tx_eoo(self.h, out.ctypes.data)
        np.copyto(tx_out, out)

    def close(self):
        if self.h:
            self.lib.rade_close(self.h); self.h = None
        if self.batch:
            self.batch.close(); self.batch = None


class radae_rx:
    """radae_rxe.radae_rx.  With bypass_dec the frames handed back are the 3 x 80 equalised latents z_hat instead of
    features (radae_rxe.py:318), as src/rade_api.c:476-506 uses it in front of its own C decoder.  The library has by then
    already run that same decoder (bit-identical to the reference's C path) on the device for the unique-word bits, so
    sum_uw_errors() from the caller is accepted and ignored instead of counted twice."""

    def __init__(self, model_name=None, latent_dim=80, auxdata=True, bottleneck=3, bpf_en=True, v=2,
                 disable_unsync=False, foff_err=0, bypass_dec=False, eoo_data_test=False, reset_decoder_on_sync=None):
        if latent_dim != 80 or not auxdata or bottleneck != 3 or not bpf_en:
            raise NotImplementedError("libradae_b200 implements the RADE V1 waveform: latent_dim=80, auxdata, bottleneck=3, bpf_en")
        if disable_unsync:
            raise NotImplementedError("disable_unsync is a reference test mode that rade_api.h does not expose")
        if foff_err not in (0, 0.0, 10, 10.0):
            raise NotImplementedError("foff_err: rade_api.h only exposes RADE_FOFF_TEST (= 10 Hz, src/rade_api.c:263-264)")
        # The reference class owns its decoder and clears its state on every candidate -> sync transition (radae_rxe.py:263);
        # only the C API with RADE_USE_C_DECODER (bypass_dec + rade_dec.c outside the class, src/rade_api.c:494-506) never does.
        # reset_decoder_on_sync: None = follow the reference (reset unless bypass_dec), or force either behaviour.
        if reset_decoder_on_sync is None:
            reset_decoder_on_sync = not bypass_dec
        flags = _FLAGS if not reset_decoder_on_sync else (_FLAGS & ~capi.RADE_USE_C_DECODER)
        if foff_err:
            flags |= capi.RADE_FOFF_TEST
        self.bypass_dec = bool(bypass_dec)
        self.h = self.batch = None
        if self.bypass_dec:
            from .batch import RadeBatch
            self.batch = RadeBatch(1, flags=flags, weights=_weights_of(model_name))
            self._nin, self._sync, self._snr = capi.NMF, False, 0
        else:
            self.lib, self.h = _open(model_name, flags)
        self.n_floats_out = capi.NZMF * capi.LATENT if self.bypass_dec else capi.NFEAT
        self._eoo = np.zeros(capi.NEOO_BITS, np.float32)

    def get_n_features_out(self): return capi.NFEAT
    def get_n_eoo_features_out(self): return capi.NEOO_BITS // 2
    def get_n_floats_out(self): return self.n_floats_out
    def get_nin_max(self): return capi.NIN_MAX
    def get_nin(self): return self._nin if self.batch else self.lib.rade_nin(self.h)
    def get_sync(self): return self._sync if self.batch else bool(self.lib.rade_sync(self.h))
    def get_snrdB_3k_est(self): return self._snr if self.batch else self.lib.rade_snrdB_3k_est(self.h)
    def get_Neoo_bits(self): return capi.NEOO_BITS
    def sum_uw_errors(self, new_uw_errors): pass          # see the class docstring

    def do_radae_rx(self, buffer_complex, floats_out):
        """returns valid_output | endofover << 1 (radae_rxe.py:330); on EOO floats_out starts with the 180 soft bits"""
        nin = self.get_nin()
        x = np.zeros(self.get_nin_max(), np.complex64)
        x[:nin] = np.asarray(buffer_complex, np.complex64)[:nin]
        if self.batch:
            _, ret, eoo = self.batch.rx(x[None, :])
            st = self.batch.rx_status()[0]
            self._nin, self._sync, self._snr = int(st.nin), st.state == 2, int(st.snrdB_3k_est)
            valid, has_eoo = int(ret[0]) & 1, (int(ret[0]) >> 1) & 1
            if valid:
                np.copyto(floats_out, self.batch.rx_z_hat()[0])
            if has_eoo:
                floats_out[:] = 0
                floats_out[:eoo.shape[1]] = eoo[0]
            return valid | (has_eoo << 1)
        feats = np.zeros(self.n_floats_out, np.float32)
        has_eoo = C.c_int(0)
        n = self.lib.rade_rx(self.h, feats.ctypes.data, C.byref(has_eoo), self._eoo.ctypes.data, x.ctypes.data)
        if n:
            np.copyto(floats_out, feats)
        if has_eoo.value:
            floats_out[:] = 0
            floats_out[:self._eoo.size] = self._eoo
        return (1 if n else 0) | (has_eoo.value << 1)

    def close(self):
        if self.h:
            self.lib.rade_close(self.h); self.h = None
        if self.batch:
            self.batch.close(); self.batch = None
