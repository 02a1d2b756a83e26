"""ctypes front-ends for the two CPU core-codec oracles (TEST INFRASTRUCTURE).

* CoreOraclePort — oracle/core_oracle.c, our restatement ("port"), built from source anywhere gcc exists.
* CoreOracleRef  — oracle/_ref/librade_ref_{int8,f32}.so: the reference's own src/rade_enc.c, rade_dec.c and
  weight tables compiled against oracle/nnet_shim (built where /root/reference exists; the .so travels).
"""
import ctypes, os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
_c = ctypes
_P = _c.c_void_p


def _rdw_default():
    return os.path.join(REPO, "radae_b200", "weights", "model19_check3.rdw")


def ensure_built():
    import importlib.util
    spec = importlib.util.spec_from_file_location("oracle_build", os.path.join(HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    mod.build_all()


class CoreOraclePort:
    """Stateful multi-stream core encoder/decoder; features [S,T,84] <-> z [S,T,80]."""
    kind = "port"

    def __init__(self, rdw_path=None, n_streams=1, bottleneck=3):
        """rdw_path: default model19_check3; an RDW of a model without the aux symbol (model05) makes the rows 80 wide;
        bottleneck 1 applies tanh to z (src/rade_enc.c:107-113)"""
        path = os.path.join(HERE, "_ref", "libcore_oracle.so")
        if not os.path.exists(path):
            ensure_built()
        self.lib = lib = _c.CDLL(path)
        lib.oracle_core_open.restype = _P
        lib.oracle_core_open.argtypes = [_c.c_char_p]
        lib.oracle_core_encode.argtypes = [_P, _P, _c.c_int, _c.c_int, _P, _P, _P, _c.c_int]
        lib.oracle_core_decode.argtypes = [_P, _P, _c.c_int, _c.c_int, _P, _P, _P, _c.c_int]
        self.h = lib.oracle_core_open((rdw_path or _rdw_default()).encode())
        if not self.h:
            raise RuntimeError("oracle_core_open failed")
        lib.oracle_core_set_bottleneck.argtypes = [_P, _c.c_int]
        lib.oracle_core_enc_in.argtypes = [_P]; lib.oracle_core_dec_out.argtypes = [_P]
        lib.oracle_core_set_bottleneck(self.h, bottleneck)
        self.in_dim, self.out_dim = lib.oracle_core_enc_in(self.h), lib.oracle_core_dec_out(self.h)
        self.n = n_streams
        self.reset()

    def reset(self):
        self.enc_state = np.zeros((self.n, self.lib.oracle_enc_state_floats()), np.float32)
        self.dec_state = np.zeros((self.n, self.lib.oracle_dec_state_floats()), np.float32)

    def encode(self, features, nthreads=1, want_cat=False):
        f = np.ascontiguousarray(features, np.float32)
        S, T, _ = f.shape
        assert S == self.n and f.shape[2] == self.in_dim
        z = np.zeros((S, T, 80), np.float32)
        cat = np.zeros((S, T, 864), np.float32) if want_cat else None
        self.lib.oracle_core_encode(self.h, self.enc_state.ctypes.data, S, T, f.ctypes.data, z.ctypes.data,
                                    cat.ctypes.data if want_cat else None, nthreads)
        return (z, cat) if want_cat else z

    def decode(self, z, nthreads=1, want_cat=False):
        zz = np.ascontiguousarray(z, np.float32)
        S, T, _ = zz.shape
        assert S == self.n and zz.shape[2] == 80
        f = np.zeros((S, T, self.out_dim), np.float32)
        cat = np.zeros((S, T, 736), np.float32) if want_cat else None
        self.lib.oracle_core_decode(self.h, self.dec_state.ctypes.data, S, T, zz.ctypes.data, f.ctypes.data,
                                    cat.ctypes.data if want_cat else None, nthreads)
        return (f, cat) if want_cat else f


class CoreOracleRef:
    """The reference's own C core codec (rade_core_encoder/decoder) behind the same interface."""
    kind = "reference"

    @staticmethod
    def available(variant="int8"):
        return os.path.exists(os.path.join(HERE, "_ref", f"librade_ref_{variant}.so"))

    def __init__(self, variant="int8", n_streams=1, blob=None, input_dim=84, output_dim=84, bottleneck=3):
        """blob: a DNNw weight file's bytes (e.g. the reference's bin/model05.bin, loaded the way src/test_rade_enc.c:50-66
        does) instead of the compiled-in model19_check3 tables; input_dim / output_dim = 4 x (20 features [+ 1 aux symbol]);
        bottleneck 1 applies tanh to z (src/rade_enc.c:107-113)"""
        self.in_dim, self.out_dim, self.bottleneck = input_dim, output_dim, bottleneck
        self.lib = lib = _c.CDLL(os.path.join(HERE, "_ref", f"librade_ref_{variant}.so"))
        lib.ref_core_open.restype = _P
        lib.ref_core_open.argtypes = [_c.c_char_p, _c.c_int, _c.c_int, _c.c_int]
        lib.ref_core_encode.argtypes = [_P, _P, _c.c_int, _c.c_int, _P, _c.c_int, _P, _c.c_int, _c.c_int]
        lib.ref_core_decode.argtypes = [_P, _P, _c.c_int, _c.c_int, _P, _P, _c.c_int, _c.c_int]
        lib.ref_core_max_abs_acc.restype = _c.c_double
        self._blob = bytes(blob) if blob is not None else None
        self.h = lib.ref_core_open(self._blob, len(self._blob) if self._blob else 0, input_dim, output_dim)
        if not self.h:
            raise RuntimeError("ref_core_open failed")
        self.n = n_streams
        self.reset()

    def reset(self):
        self.enc_state = _c.create_string_buffer(self.lib.ref_enc_state_size() * self.n)
        self.dec_state = _c.create_string_buffer(self.lib.ref_dec_state_size() * self.n)
        self.lib.ref_enc_state_init(self.enc_state, self.n)
        self.lib.ref_dec_state_init(self.dec_state, self.n)

    def encode(self, features, nthreads=1):
        f = np.ascontiguousarray(features, np.float32)
        S, T, _ = f.shape
        z = np.zeros((S, T, 80), np.float32)
        assert f.shape[2] == self.in_dim
        self.lib.ref_core_encode(self.h, self.enc_state, S, T, f.ctypes.data, self.in_dim, z.ctypes.data, self.bottleneck, nthreads)
        return z

    def decode(self, z, nthreads=1):
        zz = np.ascontiguousarray(z, np.float32)
        S, T, _ = zz.shape
        f = np.zeros((S, T, self.out_dim), np.float32)
        self.lib.ref_core_decode(self.h, self.dec_state, S, T, zz.ctypes.data, f.ctypes.data, self.out_dim, nthreads)
        return f

    def max_abs_acc(self, reset=False):
        return self.lib.ref_core_max_abs_acc(int(reset))


def synth_features(n_streams, n_frames, seed=1234):
    """Synthetic vocoder features, SURVEY.md §8(d): AR(1) x <- 0.9x + 0.1*sigma*N(0,1),
    sigma = [4, 1 x17, 0.5, 0.3]; returns [S, n_frames, 36] with dims 20..35 = 0."""
    sig = np.array([4.0] + [1.0] * 17 + [0.5, 0.3])
    out = np.zeros((n_streams, n_frames, 36), np.float32)
    for s in range(n_streams):
        rng = np.random.default_rng(seed + s)
        noise = rng.standard_normal((n_frames, 20))
        x = np.zeros(20)
        for t in range(n_frames):
            x = 0.9 * x + 0.1 * sig * noise[t]
            out[s, t, :20] = x
    return out


def pack_enc_input(features36):
    """[S, 4*T, 36] -> [S, T, 84]: 20 used features + aux = -1 per 10 ms vector (src/rade_api.c:426-432)."""
    S, F, _ = features36.shape
    x = np.concatenate([features36[:, :, :20], -np.ones((S, F, 1), np.float32)], axis=2)
    return np.ascontiguousarray(x.reshape(S, F // 4, 84))
