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


class radae_tx:
    def __init__(self, model_name=None, latent_dim=80, auxdata=True, bottleneck=3, txbpf_en=False, bypass_enc=False):
        if latent_dim != 80 or not auxdata or bottleneck != 3:
            raise NotImplementedError("libradae_b200 implements the RADE V1 waveform: latent_dim=80, auxdata, bottleneck=3")
        if txbpf_en:
            raise NotImplementedError("txbpf_en is not on the rade_api.h path (src/rade_api.c:148-150 never sets it)")
        if bypass_enc:
            raise NotImplementedError("bypass_enc: the core encoder runs on the device inside rade_tx")
        self.lib, self.h = _open(model_name, capi.RADE_USE_C_ENCODER | capi.RADE_USE_C_DECODER | capi.RADE_VERBOSE_0)
        self.n_floats_in = self.lib.rade_n_features_in_out(self.h)
        self.Nmf = self.lib.rade_n_tx_out(self.h)
        self.Neoo = self.lib.rade_n_tx_eoo_out(self.h)

    def get_n_features_in(self): return self.lib.rade_n_features_in_out(self.h)
    def get_n_floats_in(self): return self.n_floats_in
    def get_Nmf(self): return self.Nmf
    def get_Neoo(self): return self.Neoo
    def get_Neoo_bits(self): return self.lib.rade_n_eoo_bits(self.h)

    def set_eoo_bits(self, eoo_bits):
        b = np.ascontiguousarray(eoo_bits, np.float32)
        assert b.size == self.get_Neoo_bits()
        self.lib.rade_tx_set_eoo_bits(self.h, b.ctypes.data)

    def do_radae_tx(self, buffer_f32, tx_out):
        f = np.ascontiguousarray(buffer_f32, np.float32)
        assert f.size == self.n_floats_in and tx_out.dtype == np.complex64 and tx_out.size == self.Nmf
        out = np.empty(self.Nmf, np.complex64)
        self.lib.rade_tx(self.h, out.ctypes.data, f.ctypes.data)
        np.copyto(tx_out, out)

    def do_eoo(self, tx_out):
        out = np.empty(self.Neoo, np.complex64)
        self.lib.rade_tx_eoo(self.h, out.ctypes.data)
        np.copyto(tx_out, out)

    def close(self):
        if self.h:
            self.lib.rade_close(self.h); self.h = None


class radae_rx:
    def __init__(self, model_name=None, latent_dim=80, auxdata=True, bottleneck=3, bpf_en=True, v=2,
                 disable_unsync=False, foff_err=0, bypass_dec=False, eoo_data_test=False):
        if latent_dim != 80 or not auxdata or bottleneck != 3 or not bpf_en:
            raise NotImplementedError("libradae_b200 implements the RADE V1 waveform: latent_dim=80, auxdata, bottleneck=3, bpf_en")
        if disable_unsync:
            raise NotImplementedError("disable_unsync is a reference test mode that rade_api.h does not expose")
        if foff_err not in (0, 0.0, 10, 10.0):
            raise NotImplementedError("foff_err: rade_api.h only exposes RADE_FOFF_TEST (= 10 Hz, src/rade_api.c:263-264)")
        flags = capi.RADE_USE_C_ENCODER | capi.RADE_USE_C_DECODER | capi.RADE_VERBOSE_0
        if foff_err:
            flags |= capi.RADE_FOFF_TEST
        self.bypass_dec = bypass_dec
        self.lib, self.h = _open(model_name, flags)
        self.n_floats_out = self.lib.rade_n_features_in_out(self.h)
        self._eoo = np.zeros(self.lib.rade_n_eoo_bits(self.h), np.float32)

    def get_n_features_out(self): return self.lib.rade_n_features_in_out(self.h)
    def get_n_eoo_features_out(self): return self.lib.rade_n_eoo_bits(self.h) // 2
    def get_n_floats_out(self): return self.n_floats_out
    def get_nin_max(self): return self.lib.rade_nin_max(self.h)
    def get_nin(self): return self.lib.rade_nin(self.h)
    def get_sync(self): return bool(self.lib.rade_sync(self.h))
    def get_snrdB_3k_est(self): return self.lib.rade_snrdB_3k_est(self.h)
    def get_Neoo_bits(self): return self.lib.rade_n_eoo_bits(self.h)

    def do_radae_rx(self, buffer_complex, floats_out):
        """returns valid_output | endofover << 1 (radae_rxe.py:330); on EOO floats_out starts with the 180 soft bits"""
        nin = self.get_nin()
        x = np.zeros(self.get_nin_max(), np.complex64)
        x[:nin] = np.asarray(buffer_complex, np.complex64)[:nin]
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
