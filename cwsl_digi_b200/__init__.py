"""cwsl_digi_b200 -- B200-native (sm_100a) receive front-end of CWSL_DIGI.

The product is the C-ABI shared library ``libcwsl_b200.so`` (include/cwsl_b200.h) built from
``cwsl_digi_b200/csrc`` and the C++ host classes in ``cwsl_digi_b200/host``. This Python package
is only a thin ctypes binding used by the tests and bench.py; it contains no compute and no
CPU fallback: if the library is missing it raises.
"""
from .capi import (CwslError, Receiver, HostBuffer, MODE_EXACT, MODE_FAST, MODE_STFT, af_size, accepted_blocks,  # noqa: F401
                   build_tables, stft_tables, stft_channel, stft_items, ssbd_params, device_count, measure_fp32_peak, lib, lib_path, build_library)

__all__ = ["CwslError", "Receiver", "HostBuffer", "MODE_EXACT", "MODE_FAST", "MODE_STFT", "af_size", "accepted_blocks", "build_tables", "stft_tables", "stft_channel", "stft_items",
           "ssbd_params", "device_count", "measure_fp32_peak", "lib", "lib_path", "build_library"]
