"""CPU tests of the C-ABI library: it loads, exports every symbol include/cwsl_b200.h declares,
its host-only entry points agree with the reference, and it fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle.oracle import af_size as oracle_af_size

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cwsl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cwsl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cw):
    L = C.CDLL(cw.lib_path())
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/cwsl_b200.h but not exported"


def test_binding_covers_header(cw):
    from cwsl_digi_b200 import capi
    assert sorted(capi.SYMBOLS) == _declared_symbols()


def test_abi_version(cw):
    assert cw.lib().cwsl_abi_version() == 3


@pytest.mark.parametrize("fs", [192000, 96000, 48000])
def test_ssbd_params_match_reference(cw, ref, fs):
    p = cw.ssbd_params(fs)
    g = ref.getters(fs)
    for k in g:
        assert p[k] == g[k]
    assert p["NumWS"] == 32 and p["FiltOrder"] == 32 * p["BlockSize"] and p["BlockSize"] == fs // 12000


def test_ssbd_params_reject_bad_rate(cw):
    for fs in (0, 12000, 100000, 191999):       # SSBD.hpp:54: Fs/B must be an even integer >= 4
        with pytest.raises(cw.CwslError):
            cw.ssbd_params(fs)


@pytest.mark.parametrize("fs", [192000, 96000, 48000])
def test_host_tables_bit_identical_to_reference(cw, ref, fs):
    rng = np.random.default_rng(fs)
    lim = fs // 2
    freqs = [-lim, lim - 6000, 0, 1, -1, -26000 * fs // 192000] + list(rng.integers(-lim, lim - 6000, 40))
    for f in freqs:
        a = ref.tables(fs, int(f))
        b = cw.build_tables(fs, int(f))
        for k in ("filter", "tone", "phase_inc"):
            assert np.array_equal(_bits(a[k]), _bits(b[k])), (fs, f, k)


def test_host_tables_lsb(cw, ref):
    a, b = ref.tables(192000, 26000, is_usb=False), cw.build_tables(192000, 26000, is_usb=False)
    for k in ("filter", "tone", "phase_inc"):
        assert np.array_equal(_bits(a[k]), _bits(b[k]))


def test_out_of_band_rejected_like_tune(cw):
    for f in (-96001, 90001, 200000):
        with pytest.raises(cw.CwslError) as e:
            cw.build_tables(192000, f)
        assert e.value.code == -1
    cw.build_tables(192000, -96000)
    cw.build_tables(192000, 90000)


@pytest.mark.parametrize("period", [7.5, 15.0, 30.0, 60.0, 120.0, 300.0, 900.0, 1800.0])
def test_af_size(cw, period):
    assert cw.af_size(period) == oracle_af_size(period) == int(12000 * (period + 5))


@pytest.mark.parametrize("n,iq_len,fs,afs", [(100, 2048, 192000, 240000), (2000, 2048, 192000, 240000),
                                             (5, 4096, 48000, 6000), (1, 512, 192000, 600), (3, 512, 192000, 512)])
def test_accepted_blocks_matches_oracle(cw, port, n, iq_len, fs, afs):
    assert cw.accepted_blocks(n, iq_len, fs, afs) == port.accepted_blocks(n * iq_len, iq_len, fs // 12000, afs)


def test_no_gpu_fails_loudly(cw):
    if cw.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(cw.CwslError) as e:
        cw.Receiver(0, 192000, 2048)
    assert "no CUDA device" in str(e.value)
    with pytest.raises(cw.CwslError):
        cw.measure_fp32_peak(0)


def test_product_does_not_import_oracle():
    """The product path may never route through the oracle (or /root/reference)."""
    pkg = os.path.join(ROOT, "cwsl_digi_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cpp", ".hpp", ".h", ".inc")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn
                assert "liboracle" not in txt and "libcwsl_ref" not in txt, fn
                assert '#include "SSBD.hpp"' not in txt and "/root/reference" not in txt, fn


# ---- property tests (hypothesis): host arithmetic of the C ABI against the oracle restatement ----------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=150, deadline=None)
@given(n=st.integers(0, 4000), k=st.integers(1, 16), fs=st.sampled_from([48000, 96000, 192000]),
       afs=st.integers(2, 400000))
def test_accepted_blocks_property(cw, port, n, k, fs, afs):
    iq_len = k * (2 * fs // 6000)                      # any multiple of SSBD::GetInSize()
    got = cw.accepted_blocks(n, iq_len, fs, afs)
    assert got == port.accepted_blocks(n * iq_len, iq_len, fs // 12000, afs)
    assert got <= n
    if got < n:                                        # the guard tripped: one more block would not have fitted
        assert got * (iq_len // (fs // 12000)) + iq_len > afs - 1


@settings(max_examples=60, deadline=None)
@given(fs=st.sampled_from([48000, 96000, 192000]), frac=st.floats(-0.5, 0.5), usb=st.booleans())
def test_tables_property(cw, ref, fs, frac, usb):
    f = int(frac * fs)
    legal = abs(f) <= fs // 2 and abs(f + (6000 if usb else -6000)) <= fs // 2      # SSBD.hpp:100-103
    if legal:
        a, b = ref.tables(fs, f, is_usb=usb), cw.build_tables(fs, f, is_usb=usb)
        for key in ("tone", "phase_inc"):
            assert np.array_equal(_bits(a[key]), _bits(b[key])), (fs, f, usb, key)
    else:
        with pytest.raises(cw.CwslError):
            cw.build_tables(fs, f, is_usb=usb)
        with pytest.raises(ValueError):
            ref.tables(fs, f, is_usb=usb)
