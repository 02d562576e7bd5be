"""ctypes binding of include/cwsl_b200.h (one Python method per C entry point)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MODE_EXACT, MODE_FAST, MODE_STFT = 0, 1, 2

# every symbol include/cwsl_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "cwsl_abi_version", "cwsl_last_error", "cwsl_device_count", "cwsl_ssbd_params", "cwsl_build_tables",
    "cwsl_stft_tables", "cwsl_stft_channel", "cwsl_af_size", "cwsl_accepted_blocks", "cwsl_rx_create", "cwsl_rx_destroy", "cwsl_rx_set_mode",
    "cwsl_rx_add_group", "cwsl_rx_add_channel", "cwsl_rx_num_groups", "cwsl_rx_num_channels",
    "cwsl_rx_group_af_size", "cwsl_rx_push_iq", "cwsl_rx_push_iq_device", "cwsl_rx_bind_device_iq",
    "cwsl_rx_process", "cwsl_rx_end_slot", "cwsl_rx_end_slot_packed", "cwsl_rx_device_audio", "cwsl_rx_copy_device_audio", "cwsl_rx_read_float_audio",
    "cwsl_rx_channel_stats", "cwsl_rx_synchronize", "cwsl_rx_wait_output", "cwsl_rx_stream", "cwsl_rx_set_stream", "cwsl_rx_enable_timing",
    "cwsl_rx_kernel_times", "cwsl_measure_fp32_peak", "cwsl_host_alloc", "cwsl_host_free",
    "cwsl_rx_set_stft_guard", "cwsl_rx_remove_channel", "cwsl_rx_kernel_times_ex", "cwsl_rx_guard_stats",
    "cwsl_rx_push_fence", "cwsl_rx_wait_fence", "cwsl_rx_join_output", "cwsl_stft_items",
]


class CwslError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cwsl_b200 error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    # CWSL_B200_LIB selects another build of the same library (A/B experiments); default: the in-tree build
    return os.environ.get("CWSL_B200_LIB") or os.path.join(_HERE, "libcwsl_b200.so")


def build_library(quiet: bool = True) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> cwsl_digi_b200/libcwsl_b200.so"""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)
    return lib_path()


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise CwslError(-2, f"{path} is missing: build it with `make -C cwsl_digi_b200/csrc` "
                            "(this package has no CPU path)")
    L = C.CDLL(path)
    vp, u32, i32, sz = C.c_void_p, C.c_uint32, C.c_int32, C.c_size_t
    L.cwsl_abi_version.restype = C.c_int
    L.cwsl_last_error.restype = C.c_char_p
    L.cwsl_device_count.restype = C.c_int
    L.cwsl_ssbd_params.argtypes = [u32, C.POINTER(u32)]
    L.cwsl_build_tables.argtypes = [u32, i32, C.c_int, vp, vp, vp]
    L.cwsl_stft_tables.argtypes = [u32, vp, vp]
    L.cwsl_stft_channel.argtypes = [u32, i32, C.c_int, C.POINTER(C.c_int32), vp, vp]
    L.cwsl_af_size.restype = sz
    L.cwsl_af_size.argtypes = [C.c_double]
    L.cwsl_accepted_blocks.restype = sz
    L.cwsl_accepted_blocks.argtypes = [sz, u32, u32, sz]
    L.cwsl_rx_create.restype = vp
    L.cwsl_rx_create.argtypes = [C.c_int, u32, u32, C.c_double]
    L.cwsl_rx_destroy.restype = None
    L.cwsl_rx_destroy.argtypes = [vp]
    L.cwsl_rx_set_mode.argtypes = [vp, C.c_int]
    L.cwsl_rx_add_group.argtypes = [vp, C.c_double]
    L.cwsl_rx_add_channel.argtypes = [vp, C.c_int, i32, C.c_int, C.c_float]
    L.cwsl_rx_num_groups.argtypes = [vp]
    L.cwsl_rx_num_channels.argtypes = [vp, C.c_int]
    L.cwsl_rx_group_af_size.restype = sz
    L.cwsl_rx_group_af_size.argtypes = [vp, C.c_int]
    L.cwsl_rx_push_iq.argtypes = [vp, vp, sz]
    L.cwsl_rx_push_iq_device.argtypes = [vp, vp, sz]
    L.cwsl_rx_bind_device_iq.argtypes = [vp, vp, sz]
    L.cwsl_rx_process.argtypes = [vp, C.c_int]
    L.cwsl_rx_end_slot.argtypes = [vp, C.c_int, vp, C.POINTER(sz)]
    L.cwsl_rx_end_slot_packed.argtypes = [vp, C.c_int, vp, C.POINTER(sz)]
    L.cwsl_rx_device_audio.restype = vp
    L.cwsl_rx_device_audio.argtypes = [vp, C.c_int]
    L.cwsl_rx_copy_device_audio.argtypes = [vp, C.c_int, C.c_int, vp]
    L.cwsl_rx_read_float_audio.argtypes = [vp, C.c_int, C.c_int, vp]
    L.cwsl_rx_channel_stats.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.cwsl_rx_synchronize.argtypes = [vp]
    L.cwsl_rx_wait_output.argtypes = [vp]
    L.cwsl_rx_stream.restype = vp
    L.cwsl_rx_stream.argtypes = [vp]
    L.cwsl_rx_set_stream.argtypes = [vp, vp]
    L.cwsl_rx_enable_timing.argtypes = [vp, C.c_int]
    L.cwsl_rx_kernel_times.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int)]
    L.cwsl_rx_set_stft_guard.argtypes = [vp, C.c_double]
    L.cwsl_rx_remove_channel.argtypes = [vp, C.c_int, C.c_int]
    L.cwsl_rx_kernel_times_ex.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.cwsl_rx_guard_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.cwsl_rx_join_output.argtypes = [vp]
    L.cwsl_stft_items.argtypes = [C.c_uint32, vp, vp, C.c_uint32, vp, vp, vp, C.POINTER(C.c_uint32)]
    L.cwsl_rx_push_fence.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.cwsl_rx_wait_fence.argtypes = [vp, C.c_uint64]
    L.cwsl_host_alloc.restype = vp
    L.cwsl_host_alloc.argtypes = [sz]
    L.cwsl_host_free.restype = None
    L.cwsl_host_free.argtypes = [vp]
    L.cwsl_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    _lib = L
    return L


def _check(rc: int) -> int:
    if rc < 0:
        raise CwslError(rc, lib().cwsl_last_error().decode())
    return rc


def device_count() -> int:
    return lib().cwsl_device_count()


def af_size(period_s: float) -> int:
    return lib().cwsl_af_size(period_s)


def accepted_blocks(n_iq_blocks: int, iq_len: int, fs: int, afsize: int) -> int:
    return lib().cwsl_accepted_blocks(n_iq_blocks, iq_len, fs, afsize)


def ssbd_params(fs: int) -> dict:
    out = (C.c_uint32 * 9)()
    _check(lib().cwsl_ssbd_params(fs, out))
    keys = ["InRate", "OutRate", "InSize", "OutSize", "Bandwidth", "Delay", "FiltOrder", "BlockSize", "NumWS"]
    return dict(zip(keys, list(out)))


def build_tables(fs: int, demod_freq: int, is_usb: bool = True) -> dict:
    p = ssbd_params(fs)
    filt = np.zeros(p["FiltOrder"], np.float32)
    tone = np.zeros(2 * p["BlockSize"], np.float32)
    pinc = np.zeros(2, np.float32)
    _check(lib().cwsl_build_tables(fs, demod_freq, int(is_usb), filt.ctypes.data, tone.ctypes.data, pinc.ctypes.data))
    return dict(filter=filt, tone=tone, phase_inc=pinc)


def stft_tables(fs: int = 192000) -> dict:
    """Host-side shared tables of CWSL_MODE_STFT: deconvolved window [L = FiltOrder], FFT twiddles complex64
    [N/32, 32] with N = 2 L (1024 / 512 / 256 bins at 192 / 96 / 48 kHz)."""
    n_taps = ssbd_params(fs)["FiltOrder"]
    a = 2 * n_taps // 32
    w = np.zeros(n_taps, np.float32)
    tw = np.zeros(2 * a * 32, np.float32)
    _check(lib().cwsl_stft_tables(fs, w.ctypes.data, tw.ctypes.data))
    return dict(window=w, twiddle=tw.view(np.complex64).reshape(a, 32), grid=2 * n_taps)


def stft_channel(fs: int, demod_freq: int, is_usb: bool = True) -> dict:
    """Per-channel constants of CWSL_MODE_STFT: first stencil bin, 8 weights, rotation e^{-i 240 w}."""
    q0 = C.c_int32()
    wgt = np.zeros(8, np.float32)
    rot = np.zeros(2, np.float32)
    _check(lib().cwsl_stft_channel(fs, demod_freq, int(is_usb), C.byref(q0), wgt.ctypes.data, rot.ctypes.data))
    return dict(q0=q0.value, wgt=wgt, rot=complex(rot[0], rot[1]))


def stft_items(fs: int, demod_freqs, is_usb=None) -> dict:
    """Work items of the channelizer kernel for a channel set: first_bin [n_items], channels [n_items, 4] (-1: empty
    slot), weights [n_items, 4, 9]; member j of an item reads bins first_bin + (0, 0, 1, 2)[j] ... + 8."""
    f = np.ascontiguousarray(demod_freqs, np.int32)
    u = np.ascontiguousarray(np.ones(f.size, np.int32) if is_usb is None else np.asarray(is_usb, np.int32))
    first = np.zeros(f.size, np.int32)
    chans = np.full((f.size, 4), -1, np.int32)
    w = np.zeros((f.size, 4, 9), np.float32)
    n = C.c_uint32()
    _check(lib().cwsl_stft_items(fs, f.ctypes.data, u.ctypes.data, f.size, first.ctypes.data, chans.ctypes.data,
                                 w.ctypes.data, C.byref(n)))
    return dict(first_bin=first[:n.value], channels=chans[:n.value], weights=w[:n.value], shift=np.array([0, 0, 1, 2]))


def measure_fp32_peak(device: int = 0) -> dict:
    a, b = C.c_float(), C.c_float()
    _check(lib().cwsl_measure_fp32_peak(device, C.byref(a), C.byref(b)))
    return dict(ffma_tflops=a.value, ffma2_tflops=b.value)


class HostBuffer:
    """Managed pinned hand-off buffer (cwsl_host_alloc) viewed as a numpy int16 array [rows, cols]."""

    def __init__(self, rows: int, cols: int):
        self._L = lib()
        self.nbytes = rows * cols * 2
        self.ptr = self._L.cwsl_host_alloc(self.nbytes)
        if not self.ptr:
            raise CwslError(-3, self._L.cwsl_last_error().decode())
        buf = (C.c_int16 * (rows * cols)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.int16).reshape(rows, cols)

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            self._L.cwsl_host_free(self.ptr)
            self.ptr = None

    __del__ = free


class Receiver:
    """One CWSL receiver (band) on one GPU: mirrors the C handle ``cwsl_rx_t``."""

    def __init__(self, device: int, fs: int, iq_len: int, ring_seconds: float = 0.0, mode: int = MODE_FAST):
        self._L = lib()
        self._h = self._L.cwsl_rx_create(device, fs, iq_len, ring_seconds)
        if not self._h:
            raise CwslError(-1, self._L.cwsl_last_error().decode())
        self.fs, self.iq_len, self.device = fs, iq_len, device
        self.set_mode(mode)

    def close(self):
        if getattr(self, "_h", None):
            self._L.cwsl_rx_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_mode(self, mode: int):
        _check(self._L.cwsl_rx_set_mode(self._h, mode))

    def set_stft_guard(self, db_below_band_power: float):
        """Threshold of the STFT mode's dynamic-range guard (dB below the band's mean power); 0 = off."""
        _check(self._L.cwsl_rx_set_stft_guard(self._h, float(db_below_band_power)))

    def remove_channel(self, group: int, channel: int) -> None:
        _check(self._L.cwsl_rx_remove_channel(self._h, group, channel))

    def guard_stats(self, group: int) -> dict:
        """Last finished slot: channel segments the STFT guard decided, and how many it had the FAST kernel redo."""
        a, b = C.c_uint64(), C.c_uint64()
        _check(self._L.cwsl_rx_guard_stats(self._h, group, C.byref(a), C.byref(b)))
        return dict(decided=a.value, redone=b.value)

    def enable_timing(self, on: bool = True):
        _check(self._L.cwsl_rx_enable_timing(self._h, int(on)))

    def add_group(self, period_s: float) -> int:
        return _check(self._L.cwsl_rx_add_group(self._h, period_s))

    def add_channel(self, group: int, demod_freq: int, scale: float, is_usb: bool = True) -> int:
        return _check(self._L.cwsl_rx_add_channel(self._h, group, int(demod_freq), int(is_usb), scale))

    def num_channels(self, group: int) -> int:
        return _check(self._L.cwsl_rx_num_channels(self._h, group))

    def group_af_size(self, group: int) -> int:
        return self._L.cwsl_rx_group_af_size(self._h, group)

    def push_iq(self, iq) -> None:
        """iq: C-contiguous float32 numpy array (or a raw (ptr, n_blocks) tuple of pinned host memory)."""
        if isinstance(iq, tuple):
            ptr, n_blocks = iq
        else:
            iq = np.ascontiguousarray(iq, np.float32).reshape(-1)
            if iq.size % (2 * self.iq_len):
                raise ValueError("IQ length is not a whole number of blocks")
            ptr, n_blocks = iq.ctypes.data, iq.size // (2 * self.iq_len)
        _check(self._L.cwsl_rx_push_iq(self._h, ptr, n_blocks))

    def push_iq_device(self, dptr: int, n_blocks: int) -> None:
        _check(self._L.cwsl_rx_push_iq_device(self._h, dptr, n_blocks))

    def push_fence(self) -> int:
        tok = C.c_uint64()
        _check(self._L.cwsl_rx_push_fence(self._h, C.byref(tok)))
        return tok.value

    def wait_fence(self, token: int) -> None:
        _check(self._L.cwsl_rx_wait_fence(self._h, token))

    def bind_device_iq(self, dptr: int, n_blocks: int) -> None:
        _check(self._L.cwsl_rx_bind_device_iq(self._h, dptr, n_blocks))

    def process(self, group: int = -1) -> None:
        _check(self._L.cwsl_rx_process(self._h, group))

    def end_slot(self, group: int, out=None) -> int:
        """out: None (result stays on the device), a numpy int16 array [n_ch, af_size], or a raw host pointer."""
        wi = C.c_size_t()
        if out is None:
            ptr = None
        elif isinstance(out, int):
            ptr = out
        else:
            assert out.dtype == np.int16 and out.flags.c_contiguous
            ptr = out.ctypes.data
        _check(self._L.cwsl_rx_end_slot(self._h, group, ptr, C.byref(wi)))
        return wi.value

    def end_slot_packed(self, group: int, out) -> int:
        """Packed hand-off: ``out`` (a numpy int16 array or raw host pointer with room for n_ch * af_size samples)
        receives [n_ch][write_index] back to back, no zero tail. Returns write_index."""
        wi = C.c_size_t()
        ptr = out if isinstance(out, int) else out.ctypes.data
        _check(self._L.cwsl_rx_end_slot_packed(self._h, group, ptr, C.byref(wi)))
        return wi.value

    def end_slot_numpy(self, group: int):
        out = np.empty((self.num_channels(group), self.group_af_size(group)), np.int16)
        wi = self.end_slot(group, out)
        self.synchronize()
        return out, wi

    def device_audio(self, group: int) -> int:
        return self._L.cwsl_rx_device_audio(self._h, group)

    def copy_device_audio(self, group: int, channel: int, dptr: int) -> None:
        _check(self._L.cwsl_rx_copy_device_audio(self._h, group, channel, dptr))

    def read_float_audio(self, group: int, channel: int) -> np.ndarray:
        out = np.empty(self.group_af_size(group), np.float32)
        _check(self._L.cwsl_rx_read_float_audio(self._h, group, channel, out.ctypes.data))
        return out

    def channel_stats(self, group: int, channel: int):
        mx, fac = C.c_float(), C.c_float()
        _check(self._L.cwsl_rx_channel_stats(self._h, group, channel, C.byref(mx), C.byref(fac)))
        return mx.value, fac.value

    def synchronize(self) -> None:
        _check(self._L.cwsl_rx_synchronize(self._h))

    def join_output(self) -> None:
        """Device-side: later work on the receiver's stream waits for the last slot's post work (no host blocking)."""
        _check(self._L.cwsl_rx_join_output(self._h))

    def wait_output(self) -> None:
        _check(self._L.cwsl_rx_wait_output(self._h))

    def set_stream(self, cuda_stream: int) -> None:
        _check(self._L.cwsl_rx_set_stream(self._h, cuda_stream))

    def stream(self) -> int:
        return self._L.cwsl_rx_stream(self._h)

    def kernel_times(self) -> dict:
        """CUDA-event times since the last call. demod_ms: whole demodulation passes; main_ms: the demodulator
        kernel(s) alone; guard_pre_ms / guard_post_ms: the STFT guard's band-power pass and its selection + redo."""
        ms = (C.c_float * 5)()
        n = (C.c_int * 2)()
        _check(self._L.cwsl_rx_kernel_times_ex(self._h, ms, n))
        return dict(demod_ms=ms[0], quant_ms=ms[1], main_ms=ms[2], guard_pre_ms=ms[3], guard_post_ms=ms[4],
                    demod_launches=n[0], quant_launches=n[1])
