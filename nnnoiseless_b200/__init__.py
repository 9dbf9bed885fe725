"""nnnoiseless_b200 -- B200-native (sm_100a CUDA) implementation of nnnoiseless' per-frame denoise path.

This module is a thin ctypes binding over the C ABI in ``include/rnnoise.h`` (the shared library
``nnnoiseless_b200/lib/libnnnoiseless_b200.so`` built by ``nnnoiseless_b200/build.py``).  The names
mirror the reference's Rust API for the path (``src/denoise.rs``, ``src/rnn.rs``):

* :class:`RnnModel` -- ``RnnModel::{default, from_bytes}`` (+ ``from_text`` for RNNoise text models)
* :class:`DenoiseState` -- ``DenoiseState::{new, with_model, process_frame}``, one stream
* :class:`DenoiseBatch` -- N independent ``DenoiseState``s advanced by one call (additive API)

There is NO CPU fallback: if the CUDA library is missing or no GPU is visible the constructors raise.
"""
import ctypes as C
import os

import numpy as np

FRAME_SIZE = 480
NB_BANDS = 22
NB_FEATURES = 42

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NNB_LIB") or os.path.join(_HERE, "lib", "libnnnoiseless_b200.so")  # NNB_LIB: a build variant
BUILTIN_WEIGHTS_PATH = os.path.join(_HERE, "data", "weights.rnn")

# every symbol include/rnnoise.h declares (checked by tests/test_capi_symbols.py)
C_ABI_SYMBOLS = [
    "rnnoise_get_frame_size", "rnnoise_get_size", "rnnoise_init", "rnnoise_create", "rnnoise_destroy",
    "rnnoise_process_frame", "rnnoise_model_from_file", "rnnoise_model_free",
    "rnnoise_model_from_bytes", "rnnoise_model_from_text", "rnnoise_model_bytes",
    "rnnoise_batch_create", "rnnoise_batch_destroy", "rnnoise_batch_streams", "rnnoise_batch_reset",
    "rnnoise_batch_process_device", "rnnoise_batch_process_device_pcm16", "rnnoise_batch_process_device_strided",
    "rnnoise_batch_process_host",
    "rnnoise_batch_process_pcm16_host",
    "rnnoise_batch_get_taps", "rnnoise_batch_profile_step", "rnnoise_kernel_name", "rnnoise_batch_pitch_stats",
    "rnnoise_train_create", "rnnoise_train_destroy", "rnnoise_train_lanes", "rnnoise_train_set_params",
    "rnnoise_train_band_lp", "rnnoise_train_process_host", "rnnoise_train_process_device",
    "rnnoise_denoise_file", "rnnoise_denoise_files", "rnnoise_resample_host",
    "rnnoise_audio_read", "rnnoise_audio_free", "rnnoise_audio_write",
    "rnnoise_kernel_launches", "rnnoise_last_error",
]

_lib = None


class NnnoiselessError(RuntimeError):
    pass


def lib():
    """Load the CUDA shared library (fails loudly if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NnnoiselessError(
            "CUDA library %s is missing: run `python -m nnnoiseless_b200.build` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    L.rnnoise_get_frame_size.restype = ci
    L.rnnoise_get_size.restype = ci
    L.rnnoise_init.restype = ci
    L.rnnoise_init.argtypes = [vp, vp]
    L.rnnoise_create.restype = vp
    L.rnnoise_create.argtypes = [vp]
    L.rnnoise_destroy.argtypes = [vp]
    L.rnnoise_process_frame.restype = cf
    L.rnnoise_process_frame.argtypes = [vp, vp, vp]
    L.rnnoise_model_from_file.restype = vp
    L.rnnoise_model_from_file.argtypes = [vp]
    L.rnnoise_model_free.argtypes = [vp]
    L.rnnoise_model_from_bytes.restype = vp
    L.rnnoise_model_from_bytes.argtypes = [C.c_char_p, C.c_size_t]
    L.rnnoise_model_from_text.restype = vp
    L.rnnoise_model_from_text.argtypes = [C.c_char_p, C.c_size_t]
    L.rnnoise_model_bytes.restype = C.c_size_t
    L.rnnoise_model_bytes.argtypes = [vp, vp, C.c_size_t]
    L.rnnoise_batch_create.restype = vp
    L.rnnoise_batch_create.argtypes = [vp, ci, ci]
    L.rnnoise_batch_destroy.argtypes = [vp]
    L.rnnoise_batch_streams.restype = ci
    L.rnnoise_batch_streams.argtypes = [vp]
    L.rnnoise_batch_reset.restype = ci
    L.rnnoise_batch_reset.argtypes = [vp]
    L.rnnoise_batch_process_device.restype = ci
    L.rnnoise_batch_process_device.argtypes = [vp, vp, vp, vp, ci, C.c_long, C.c_long, vp]
    L.rnnoise_batch_process_device_pcm16.restype = ci
    L.rnnoise_batch_process_device_pcm16.argtypes = [vp, vp, vp, vp, ci, C.c_long, C.c_long, vp]
    L.rnnoise_batch_process_device_strided.restype = ci
    L.rnnoise_batch_process_device_strided.argtypes = [vp, vp, vp, ci, vp, ci, C.c_long, C.c_long, C.c_long, vp]
    L.rnnoise_batch_process_host.restype = ci
    L.rnnoise_batch_process_host.argtypes = [vp, vp, vp, vp, ci]
    L.rnnoise_batch_process_pcm16_host.restype = ci
    L.rnnoise_batch_process_pcm16_host.argtypes = [vp, vp, vp, vp, ci]
    L.rnnoise_batch_get_taps.restype = ci
    L.rnnoise_batch_get_taps.argtypes = [vp, vp, vp, vp, vp]
    L.rnnoise_batch_pitch_stats.restype = ci
    L.rnnoise_batch_pitch_stats.argtypes = [vp, vp]
    L.rnnoise_batch_profile_step.restype = ci
    L.rnnoise_batch_profile_step.argtypes = [vp, vp, vp, vp, C.c_long, vp, vp, ci]
    L.rnnoise_train_create.restype = vp
    L.rnnoise_train_create.argtypes = [ci, ci]
    L.rnnoise_train_destroy.argtypes = [vp]
    L.rnnoise_train_lanes.restype = ci
    L.rnnoise_train_lanes.argtypes = [vp]
    L.rnnoise_train_set_params.restype = ci
    L.rnnoise_train_set_params.argtypes = [vp, ci, ci, vp]
    L.rnnoise_train_band_lp.restype = ci
    L.rnnoise_train_band_lp.argtypes = [ci]
    L.rnnoise_train_process_host.restype = ci
    L.rnnoise_train_process_host.argtypes = [vp, vp, vp, vp, ci]
    L.rnnoise_train_process_device.restype = ci
    L.rnnoise_train_process_device.argtypes = [vp, vp, vp, vp, ci, C.c_long, C.c_long, C.c_long, C.c_long, vp]
    L.rnnoise_denoise_file.restype = ci
    L.rnnoise_denoise_file.argtypes = [C.c_char_p, C.c_char_p, vp]
    L.rnnoise_denoise_files.restype = ci
    L.rnnoise_denoise_files.argtypes = [ci, vp, vp, vp]
    L.rnnoise_resample_host.restype = C.c_long
    L.rnnoise_resample_host.argtypes = [vp, C.c_long, vp, C.c_long, ci, C.c_double, ci]
    L.rnnoise_audio_read.restype = ci
    L.rnnoise_audio_read.argtypes = [C.c_char_p, ci, ci, C.c_double, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_long),
                                     C.POINTER(ci), C.POINTER(C.c_double)]
    L.rnnoise_audio_free.argtypes = [C.POINTER(C.c_float)]
    L.rnnoise_audio_write.restype = ci
    L.rnnoise_audio_write.argtypes = [C.c_char_p, ci, vp, C.c_long, ci]
    L.rnnoise_kernel_name.restype = C.c_char_p
    L.rnnoise_kernel_name.argtypes = [ci]
    L.rnnoise_kernel_launches.restype = C.c_ulonglong
    L.rnnoise_last_error.restype = C.c_char_p
    _lib = L
    return L


def last_error() -> str:
    return lib().rnnoise_last_error().decode("utf-8", "replace")


def kernel_launches() -> int:
    return int(lib().rnnoise_kernel_launches())


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RnnModel:
    """``RnnModel`` (src/rnn.rs:55-94).  ``RnnModel()`` is the built-in model (``Default``)."""

    def __init__(self, _handle=None):
        self._h = _handle  # None = built-in (NULL at the C ABI)

    @classmethod
    def from_bytes(cls, data: bytes):
        """``RnnModel::from_bytes``: returns None for malformed bytes, like the reference's Option."""
        h = lib().rnnoise_model_from_bytes(bytes(data), len(data))
        return cls(h) if h else None

    @classmethod
    def from_text(cls, text):
        """RNNoise text format (train/convert_rnnoise.py) -> model; None if malformed."""
        if isinstance(text, str):
            text = text.encode("ascii")
        h = lib().rnnoise_model_from_text(text, len(text))
        return cls(h) if h else None

    def to_bytes(self) -> bytes:
        n = lib().rnnoise_model_bytes(self._h, None, 0)
        buf = (C.c_ubyte * n)()
        lib().rnnoise_model_bytes(self._h, buf, n)
        return bytes(buf)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.rnnoise_model_free(h)
            self._h = None


class DenoiseState:
    """``DenoiseState`` (src/denoise.rs:37-116) for ONE stream, through the legacy rnnoise_* ABI."""

    FRAME_SIZE = FRAME_SIZE

    def __init__(self, model: RnnModel = None):
        self._model = model  # borrowed by the state: keep it alive (src/capi.rs:53)
        self._h = lib().rnnoise_create(model._h if model is not None else None)
        if not self._h:
            raise NnnoiselessError("rnnoise_create failed: " + last_error())

    @classmethod
    def new(cls):
        return cls()

    @classmethod
    def with_model(cls, model: RnnModel):
        return cls(model)

    from_model = with_model

    def process_frame(self, output: np.ndarray, input: np.ndarray) -> float:
        """``process_frame(&mut self, output, input) -> f32``; panics (raises) unless both are 480 long."""
        if input.shape != (FRAME_SIZE,) or output.shape != (FRAME_SIZE,):
            raise ValueError("process_frame needs 480-sample input and output")  # assert!, src/features.rs:98
        if input.dtype != np.float32 or output.dtype != np.float32:
            raise TypeError("float32 buffers required")
        if not (input.flags.c_contiguous and output.flags.c_contiguous):
            raise ValueError("contiguous buffers required")
        return float(lib().rnnoise_process_frame(self._h, _np_ptr(output), _np_ptr(input)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.rnnoise_destroy(h)
            self._h = None


class DenoiseBatch:
    """N independent ``DenoiseState``s on one GPU, advanced together (rnnoise_batch_* in include/rnnoise.h)."""

    def __init__(self, n_streams: int, model: RnnModel = None, device: int = -1):
        self.n_streams = int(n_streams)
        self._h = lib().rnnoise_batch_create(model._h if model is not None else None, self.n_streams, int(device))
        if not self._h:
            raise NnnoiselessError("rnnoise_batch_create failed: " + last_error())

    def reset(self):
        if lib().rnnoise_batch_reset(self._h) != 0:
            raise NnnoiselessError(last_error())

    def process_host(self, x: np.ndarray, want_vad=True):
        """x: [T][B][480] float32 host array -> (out [T][B][480], vad [T][B])."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        T, B, F = x.shape
        if B != self.n_streams or F != FRAME_SIZE:
            raise ValueError("expected [T][%d][480]" % self.n_streams)
        out = np.empty_like(x)
        vad = np.empty((T, B), np.float32) if want_vad else None
        rc = lib().rnnoise_batch_process_host(self._h, _np_ptr(out), _np_ptr(x), _np_ptr(vad) if want_vad else None, T)
        if rc != 0:
            raise NnnoiselessError(last_error())
        return out, vad

    def process_pcm16_host(self, x: np.ndarray):
        """x: [T][B][480] int16 -> (out int16 [T][B][480], vad [T][B])."""
        x = np.ascontiguousarray(x, dtype=np.int16)
        T, B, F = x.shape
        if B != self.n_streams or F != FRAME_SIZE:
            raise ValueError("expected [T][%d][480]" % self.n_streams)
        out = np.empty_like(x)
        vad = np.empty((T, B), np.float32)
        rc = lib().rnnoise_batch_process_pcm16_host(self._h, _np_ptr(out), _np_ptr(x), _np_ptr(vad), T)
        if rc != 0:
            raise NnnoiselessError(last_error())
        return out, vad

    def process_device(self, out_ptr: int, in_ptr: int, vad_ptr: int, n_frames: int, stream_stride: int,
                       frame_stride: int, cuda_stream: int = 0, pcm16: bool = False):
        """Raw device pointers (e.g. torch tensors' data_ptr()); strides in samples (float32, or int16 if pcm16)."""
        fn = lib().rnnoise_batch_process_device_pcm16 if pcm16 else lib().rnnoise_batch_process_device
        rc = fn(self._h, C.c_void_p(out_ptr), C.c_void_p(in_ptr),
                                                C.c_void_p(vad_ptr) if vad_ptr else None, int(n_frames),
                                                int(stream_stride), int(frame_stride),
                                                C.c_void_p(cuda_stream) if cuda_stream else None)
        if rc != 0:
            raise NnnoiselessError(last_error())

    def profile_step(self, out_ptr: int, in_ptr: int, vad_ptr: int, stream_stride: int, cuda_stream: int = 0):
        """One frame with CUDA events between the kernels: returns {kernel name: milliseconds}."""
        ms = (C.c_float * 16)()
        n = lib().rnnoise_batch_profile_step(self._h, C.c_void_p(out_ptr), C.c_void_p(in_ptr),
                                             C.c_void_p(vad_ptr) if vad_ptr else None, int(stream_stride),
                                             C.c_void_p(cuda_stream) if cuda_stream else None, ms, 16)
        if n < 0:
            raise NnnoiselessError(last_error())
        return {lib().rnnoise_kernel_name(i).decode(): float(ms[i]) for i in range(n)}

    def pitch_stats(self):
        """Cumulative pitch-kernel certification counters: dict(coarse_exact, ladder_exact, stream_frames)."""
        out = (C.c_ulonglong * 3)()
        if lib().rnnoise_batch_pitch_stats(self._h, out) != 0:
            raise NnnoiselessError(last_error())
        return dict(coarse_exact=int(out[0]), ladder_exact=int(out[1]), stream_frames=int(out[2]))

    def taps(self):
        """Intermediates of the most recent frame: dict(pitch, silence, features, gains)."""
        B = self.n_streams
        pitch = np.empty(B, np.int32)
        silence = np.empty(B, np.int32)
        feats = np.empty((B, NB_FEATURES), np.float32)
        gains = np.empty((B, NB_BANDS), np.float32)
        rc = lib().rnnoise_batch_get_taps(self._h, _np_ptr(pitch), _np_ptr(silence), _np_ptr(feats), _np_ptr(gains))
        if rc != 0:
            raise NnnoiselessError(last_error())
        return dict(pitch=pitch, silence=silence, features=feats, gains=gains)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.rnnoise_batch_destroy(h)
            self._h = None


def shard_streams(n_streams: int, world_size: int, rank: int):
    """Contiguous block sharding of independent streams over ranks: returns (start, count)."""
    base, rem = divmod(int(n_streams), int(world_size))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)
