"""ctypes bindings for the parity oracle (TEST INFRASTRUCTURE ONLY).

* ``Port``  -> oracle/liboracle.so  : plain-C restatement (oracle/cwsl_oracle.c)
* ``Ref``   -> oracle/_ref/libcwsl_ref.so : the reference's own SSBD.hpp/LowPass.hpp compiled
  by oracle/Makefile (strict IEEE flags); ``Ref(fast=True)`` loads the -ffast-math speed build.

Nothing here reads /root/reference at run time; the prebuilt libraries travel with the repo
snapshot to the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(quiet: bool = True) -> None:
    """Compile oracle/liboracle.so and, when /root/reference is present, oracle/_ref/*.so."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def af_size(period_s: float, wave_sr: int = 12000) -> int:
    """(period+5 s)*12000, source/Instance.cpp:149."""
    return int(float(wave_sr) * float(period_s + 5))


class Port:
    """The plain-C restatement."""

    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = lib = C.CDLL(path)
        lib.oracle_slot.restype = C.c_size_t
        lib.oracle_slot.argtypes = [C.c_uint32, C.c_int32, _f32p, C.c_size_t, C.c_size_t, C.c_float,
                                    C.c_size_t, C.c_void_p, _i16p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.oracle_slot_sb.restype = C.c_size_t
        lib.oracle_slot_sb.argtypes = [C.c_uint32, C.c_int32, C.c_int, _f32p, C.c_size_t, C.c_size_t, C.c_float,
                                       C.c_size_t, C.c_void_p, _i16p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.oracle_tables_flat.restype = C.c_int
        lib.oracle_tables_flat.argtypes = [C.c_uint32, C.c_int32, C.c_int, _f32p, _f32p, _f32p,
                                           C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
        lib.oracle_phase_table.restype = None
        lib.oracle_phase_table.argtypes = [_f32p, C.c_size_t, _f32p]
        lib.oracle_accepted_blocks.restype = C.c_size_t
        lib.oracle_accepted_blocks.argtypes = [C.c_size_t] * 4

    def slot(self, fs, demod_freq, iq, iq_len, scale, afsize, want_raw=True, is_usb=True):
        iq = np.ascontiguousarray(iq, dtype=np.float32).reshape(-1)
        n_iq = iq.size // 2
        out = np.zeros(afsize, np.int16)
        raw = np.zeros(afsize, np.float32) if want_raw else None
        mx, fac = C.c_float(), C.c_float()
        wi = self.lib.oracle_slot_sb(fs, demod_freq, int(is_usb), iq, n_iq, iq_len, scale, afsize,
                                     raw.ctypes.data if want_raw else None, out, C.byref(mx), C.byref(fac))
        if wi == C.c_size_t(-1).value:
            raise ValueError("invalid tuning")
        return dict(write_index=wi, i16=out, raw=raw, max=mx.value, factor=fac.value)

    def tables(self, fs, demod_freq, is_usb=True):
        filt = np.zeros(4096, np.float32)
        tone = np.zeros(2 * 128, np.float32)
        pinc = np.zeros(2, np.float32)
        rs = C.c_float()
        dims = (C.c_uint32 * 3)()
        rc = self.lib.oracle_tables_flat(fs, demod_freq, int(is_usb), filt, tone, pinc, C.byref(rs), dims)
        if rc != 0:
            raise ValueError("invalid tuning")
        return dict(filter=filt[:dims[0]].copy(), tone=tone[:2 * dims[1]].copy(), phase_inc=pinc,
                    raw_tap_sum=rs.value, filt_order=dims[0], block_size=dims[1], num_ws=dims[2])

    def phase_table(self, phase_inc, n):
        tab = np.zeros(2 * n, np.float32)
        self.lib.oracle_phase_table(np.ascontiguousarray(phase_inc, np.float32), n, tab)
        return tab.reshape(n, 2)

    def accepted_blocks(self, n_iq, iq_len, dec_ratio, afsize):
        return self.lib.oracle_accepted_blocks(n_iq, iq_len, dec_ratio, afsize)


class Ref:
    """The reference's own headers, compiled (oracle/_ref)."""

    def __init__(self, fast: bool = False):
        name = "libcwsl_ref_fast.so" if fast else "libcwsl_ref.so"
        path = os.path.join(_HERE, "_ref", name)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (needs /root/reference at build time)")
        self.fast = fast
        self.lib = lib = C.CDLL(path)
        lib.cwsl_ref_slot.restype = C.c_size_t
        lib.cwsl_ref_slot.argtypes = [C.c_uint32, C.c_int32, _f32p, C.c_size_t, C.c_size_t, C.c_float,
                                      C.c_size_t, C.c_void_p, _i16p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.cwsl_ref_slot_sb.restype = C.c_size_t
        lib.cwsl_ref_slot_sb.argtypes = [C.c_uint32, C.c_int32, C.c_int, _f32p, C.c_size_t, C.c_size_t, C.c_float,
                                         C.c_size_t, C.c_void_p, _i16p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.cwsl_ref_tables.restype = C.c_size_t
        lib.cwsl_ref_tables.argtypes = [C.c_uint32, C.c_int32, C.c_int, _f32p, _f32p, _f32p, C.POINTER(C.c_float)]
        lib.cwsl_ref_phase_after.restype = None
        lib.cwsl_ref_phase_after.argtypes = [C.c_uint32, C.c_int32, C.c_size_t, _f32p]
        lib.cwsl_ref_getters.restype = C.c_int
        lib.cwsl_ref_getters.argtypes = [C.c_uint32, C.POINTER(C.c_size_t)]
        lib.cwsl_ref_chain_threads.restype = C.c_double
        lib.cwsl_ref_chain_threads.argtypes = [C.c_uint32, _i32p, _f32p, C.c_size_t, _f32p, C.c_size_t,
                                               C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                               C.POINTER(C.c_uint64)]
        lib.cwsl_ref_hardware_concurrency.restype = C.c_uint

    def slot(self, fs, demod_freq, iq, iq_len, scale, afsize, want_raw=True, is_usb=True):
        iq = np.ascontiguousarray(iq, dtype=np.float32).reshape(-1)
        n_iq = iq.size // 2
        out = np.zeros(afsize, np.int16)
        raw = np.zeros(afsize, np.float32) if want_raw else None
        mx, fac = C.c_float(), C.c_float()
        wi = self.lib.cwsl_ref_slot_sb(fs, demod_freq, int(is_usb), iq, n_iq, iq_len, scale, afsize,
                                       raw.ctypes.data if want_raw else None, out, C.byref(mx), C.byref(fac))
        if wi == C.c_size_t(-1).value:
            raise ValueError("invalid tuning")
        return dict(write_index=wi, i16=out, raw=raw, max=mx.value, factor=fac.value)

    def tables(self, fs, demod_freq, is_usb=True):
        filt = np.zeros(4096, np.float32)
        tone = np.zeros(2 * 128, np.float32)
        pinc = np.zeros(2, np.float32)
        rs = C.c_float()
        order = self.lib.cwsl_ref_tables(fs, demod_freq, int(is_usb), filt, tone, pinc, C.byref(rs))
        if order == 0:
            raise ValueError("invalid tuning")
        bs = fs // 6000 // 2
        return dict(filter=filt[:order].copy(), tone=tone[:2 * bs].copy(), phase_inc=pinc,
                    raw_tap_sum=rs.value, filt_order=order, block_size=bs, num_ws=order // bs)

    def phase_after(self, fs, demod_freq, n_blocks):
        p = np.zeros(2, np.float32)
        self.lib.cwsl_ref_phase_after(fs, demod_freq, n_blocks, p)
        return p

    def getters(self, fs):
        v = (C.c_size_t * 6)()
        if self.lib.cwsl_ref_getters(fs, v) != 0:
            raise ValueError("invalid Fs")
        return dict(zip(["InRate", "OutRate", "InSize", "OutSize", "Bandwidth", "Delay"], list(v)))

    def hardware_concurrency(self):
        return int(self.lib.cwsl_ref_hardware_concurrency())

    def chain_threads(self, fs, demod_freqs, scales, iq, iq_len, afsize, max_threads=0, keep=False):
        """CPU baseline: thread-per-decoder chain on shared IQ. Returns (seconds, checksum, out|None)."""
        demod_freqs = np.ascontiguousarray(demod_freqs, np.int32)
        scales = np.ascontiguousarray(scales, np.float32)
        iq = np.ascontiguousarray(iq, dtype=np.float32).reshape(-1)
        n_ch = demod_freqs.size
        out = np.zeros((n_ch, afsize), np.int16) if keep else None
        cs = C.c_uint64()
        sec = self.lib.cwsl_ref_chain_threads(fs, demod_freqs, scales, n_ch, iq, iq.size // 2, iq_len, afsize,
                                              out.ctypes.data if keep else None, max_threads, C.byref(cs))
        return sec, cs.value, out
