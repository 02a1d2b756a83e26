"""ctypes binding of libradae_b200.so — the C ABI declared in include/rade_api.h and include/rade_b200.h.

The product path: there is no fallback.  If the shared library is missing or cannot be loaded this module raises
ImportError-like RuntimeError with the build command; if no sm_100 device is present `rade_b200_open` fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libradae_b200.so")

NMF, NEOO, NIN_MAX, NFEAT, NEOO_BITS, LATENT, NZMF = 960, 1152, 1120, 432, 180, 80, 3
RADE_USE_C_ENCODER, RADE_USE_C_DECODER, RADE_FOFF_TEST, RADE_VERBOSE_0 = 1, 2, 4, 8
RADE_B200_BOTTLENECK_1 = 0x100


class RxStatus(C.Structure):
    _fields_ = [("state", C.c_int), ("nin", C.c_int), ("tmax", C.c_int), ("valid_count", C.c_int),
                ("uw_errors", C.c_int), ("synced_count", C.c_int), ("snrdB_3k_est", C.c_int),
                ("snrdB_3k_est_f", C.c_float), ("fmax", C.c_double), ("Dthresh", C.c_float),
                ("Dtmax12", C.c_float), ("Dtmax12_eoo", C.c_float)]


class ChannelCfg(C.Structure):
    _fields_ = [("EbNodB", C.c_float), ("freq_offset_hz", C.c_float), ("freq_offset_spread_hz", C.c_float),
                ("doppler_spread_hz", C.c_float), ("delay_samples", C.c_int), ("gain", C.c_float),
                ("seed", C.c_ulonglong)]


_P, _I, _F = C.c_void_p, C.c_int, C.c_float
# name -> (restype, argtypes); every symbol the two public headers declare
SIGNATURES = {
    # include/rade_api.h
    "rade_initialize": (None, []), "rade_finalize": (None, []),
    "rade_open": (_P, [C.c_char_p, _I]), "rade_close": (None, [_P]), "rade_version": (_I, []),
    "rade_n_tx_out": (_I, [_P]), "rade_n_tx_eoo_out": (_I, [_P]), "rade_nin_max": (_I, [_P]),
    "rade_n_features_in_out": (_I, [_P]), "rade_n_eoo_bits": (_I, [_P]),
    "rade_tx": (_I, [_P, _P, _P]), "rade_tx_set_eoo_bits": (None, [_P, _P]), "rade_tx_eoo": (_I, [_P, _P]),
    "rade_nin": (_I, [_P]), "rade_rx": (_I, [_P, _P, C.POINTER(_I), _P, _P]), "rade_sync": (_I, [_P]),
    "rade_freq_offset": (_F, [_P]), "rade_snrdB_3k_est": (_I, [_P]),
    # include/rade_b200.h
    "rade_b200_open": (_P, [_I, _I, _I, _P, C.c_size_t]), "rade_b200_close": (None, [_P]),
    "rade_b200_n_streams": (_I, [_P]), "rade_b200_cuda_stream": (_P, [_P]), "rade_b200_synchronize": (_I, [_P]),
    "rade_b200_launch_count": (C.c_longlong, [_P]), "rade_b200_default_weights_blob": (_P, [C.POINTER(C.c_size_t)]),
    "rade_b200_reset": (_I, [_P]),
    "rade_b200_core_dims": (_I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "rade_b200_core_encode_dev": (_I, [_P, _P, _P, _I]), "rade_b200_core_decode_dev": (_I, [_P, _P, _P, _I]),
    "rade_b200_core_encode": (_I, [_P, _P, _P, _I]), "rade_b200_core_decode": (_I, [_P, _P, _P, _I]),
    "rade_b200_tx_dev": (_I, [_P, _P, _P]), "rade_b200_tx": (_I, [_P, _P, _P]),
    "rade_b200_tx_set_eoo_bits": (_I, [_P, _P]), "rade_b200_tx_eoo": (_I, [_P, _P]),
    "rade_b200_ofdm_mod_dev": (_I, [_P, _P, _P]),
    "rade_b200_tx_z_dev": (_I, [_P, _P, _P]), "rade_b200_tx_z": (_I, [_P, _P, _P]), "rade_b200_tx_bpf_enable": (_I, [_P, _I]),
    "rade_b200_nin": (_I, [_P, _P]), "rade_b200_rx": (_I, [_P, _P, _P, _P, _P, _P]),
    "rade_b200_rx_dev": (_I, [_P, _P, _P, _P, _P, _P]), "rade_b200_nin_dev": (_P, [_P]),
    "rade_b200_rx_get_status": (_I, [_P, _P]), "rade_b200_rx_get_z_hat": (_I, [_P, _P]),
    "rade_b200_channel_apply_dev": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _F, _F, _F]),
    "rade_b200_channel_apply": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _F, _F, _F]),
    "rade_b200_channel_apply_drift": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _F, _F, _F, _F]),
    "rade_b200_channel_config": (_I, [_P, C.POINTER(ChannelCfg)]), "rade_b200_channel_dev": (_I, [_P, _P, _P]),
    "rade_b200_link_push_dev": (_I, [_P, _P]), "rade_b200_channel_link_dev": (_I, [_P, _P]),
    "rade_b200_rx_link_dev": (_I, [_P, _P, _P, _P]), "rade_b200_tx_channel_link_dev": (_I, [_P, _P]),
    "rade_b200_loopback_step_dev": (_I, [_P, _P, _P, _P, _P]),
    "rade_b200_pipeline_enable": (_I, [_P, _I]), "rade_b200_pipeline_fork": (_I, [_P]), "rade_b200_pipeline_join": (_I, [_P]), "rade_b200_link_pop_dev": (_I, [_P, _P, _P]),
    "rade_b200_profile_enable": (_I, [_P, _I]), "rade_b200_profile_n_kernels": (_I, []),
    "rade_b200_profile_kernel_name": (C.c_char_p, [_I]), "rade_b200_profile_read": (_I, [_P, _P, _P]),
    "rade_b200_channel": (_I, [_P, _P, _P]),
    "rade_b200_hostlink_open": (_P, [_P, _I]), "rade_b200_hostlink_close": (None, [_P]),
    "rade_b200_hostlink_push": (_I, [_P, _P]), "rade_b200_hostlink_rx": (_I, [_P, _P, _P, _P]),
    "rade_b200_hostlink_active": (_P, [_P]), "rade_b200_hostlink_dropped": (C.c_longlong, [_P]),
    "rade_b200_channel_hostlink": (_I, [_P, _P, _P]),
    "rade_b200_loopback_run": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "rade_b200_host_alloc": (_P, [C.c_size_t]), "rade_b200_host_free": (None, [_P]),
    "rade_b200_open_multi": (_P, [_I, C.c_ulonglong, _I, C.c_char_p, C.c_size_t]),
    "rade_b200_open_devices": (_P, [_I, C.POINTER(_I), _I, _I, C.c_char_p, C.c_size_t]),
    "rade_b200_close_multi": (None, [_P]), "rade_b200_multi_n_devices": (_I, [_P]), "rade_b200_multi_n_streams": (_I, [_P]),
    "rade_b200_multi_context": (_P, [_P, _I, C.POINTER(_I), C.POINTER(_I)]),
    "rade_b200_multi_tx": (_I, [_P, _P, _P]), "rade_b200_multi_nin": (_I, [_P, _P]),
    "rade_b200_multi_rx": (_I, [_P, _P, _P, _P, _P, _P]), "rade_b200_multi_rx_get_status": (_I, [_P, _P]),
    "rade_b200_duplex_run": (_I, [_P, _P, _P, _P, _I, _I, _P, _I, _P, _P, _P]),
    # test hook
    "rade_b200_debug_tables": (_I, [_I, _P, _I]),
    "rade_b200_timeline_begin": (_I, [_P]),
    "rade_b200_timeline_read": (_I, [_P, _P, _P, _P, _I]),
    "rade_b200_debug_codec_stream": (C.c_longlong, [_I, _I, _P, C.c_longlong, _P, _I, C.POINTER(_I), C.POINTER(_I)]),
    "rade_b200_debug_codec_program": (_I, [_I, _P, _I]),
    "rade_b200_debug_check_weights": (_I, [_P, C.c_size_t]),
    "rade_b200_debug_trace_enable": (_I, [_P, _I]),
    "rade_b200_debug_trace_read": (_I, [_P, _P, _I]),
}

_lib = None


def lib():
    """Load the CUDA library (building is a separate, explicit step: `python -m radae_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing — build it with `python -m radae_b200.build` "
                               "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError here = the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib
