"""CPU checks on the compiled sm_100a code (cuobjdump -sass): the EXACT kernels must contain no fused
multiply-add (ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2, which would silently break
bit-exactness), the FAST kernel must be built from FFMA2 with TMA bulk copies (UBLKCP) and no local-memory spills."""
import re
import shutil
import subprocess

import pytest


def _sass(cw):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", cw.lib_path()], capture_output=True, text=True, check=True).stdout
    funcs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[name].append(line.split("*/", 1)[1].strip())
    return funcs


def _ops(body):
    return [ins.split()[1].split(".")[0] if ins.startswith("@") else ins.split()[0].split(".")[0] for ins in body if ins]


def test_exact_kernels_have_no_fused_multiply_add(cw):
    funcs = _sass(cw)
    # (quantise_kernel legitimately contains FFMA: it is the Newton iteration inside the IEEE-rounded __fdiv_rn)
    exact = {k: v for k, v in funcs.items() if "demod_exact" in k or "phase_table" in k}
    assert len(exact) >= 7
    for name, body in exact.items():
        ops = _ops(body)
        assert "FFMA" not in ops and "FFMA2" not in ops, f"{name} contains a fused multiply-add"
        assert "FMUL" in ops or "FMUL2" in ops


def test_fast_kernel_is_ffma2_tma_and_spill_free(cw):
    funcs = _sass(cw)
    fast = {k: v for k, v in funcs.items() if "demod_fast_kernelILi16ELi4ELi128ELi2ELb0E" in k}
    assert len(fast) == 2        # the production instantiations (no register prefetch): direct grid + guard work list
    for name, body in fast.items():
        ops = _ops(body)
        assert ops.count("FFMA2") > 400                     # packed FP32 FMA (Blackwell)
        assert "UBLKCP" in ops                              # cp.async.bulk = TMA bulk copy engine
        assert any(o.startswith("SYNCS") for o in ops)      # mbarrier
        assert "LDL" not in ops and "STL" not in ops        # no register spills
        imm = [ins for ins in body if ins.startswith("FFMA2") and re.search(r", -?[0-9]\.[0-9e+-]+, ", ins)]
        assert len(imm) > 300                               # taps are FFMA2 immediates: no loads in the inner product
    tiled = [v for k, v in funcs.items() if "demod_exact_tiled_kernelILi16" in k][0]
    assert "UBLKCP" in _ops(tiled)


def test_channelizer_kernel_shape(cw):
    """STFT channelizer: IQ staged by the TMA bulk-copy engine behind mbarriers, packed FP32 butterflies and packed
    complex twiddle multiplies, 128-bit spectrum reads shared by the four channels of a work item, one 256-bit store
    per channel and eight hops, named-barrier hand-over between the FFT warps and the interpolation warps, registers
    moved between the roles with setmaxnreg (USETMAXREG), no spills, and few enough registers at launch
    (<= 120 x 512 threads) that one CTA of the quantise kernel fits beside it on every SM."""
    funcs = _sass(cw)
    chan = {k: v for k, v in funcs.items() if "demod_chan_kernel" in k}
    assert len(chan) == 3                               # one per receiver rate
    for name, body in chan.items():
        ops = _ops(body)
        assert ops.count("FADD2") > 100 and ops.count("FFMA2") >= 36 * 8 and ops.count("FMUL2") >= 100
        assert sum(i.startswith("LDS.128") for i in body) >= 6 * 8
        assert "UBLKCP" in ops and any(o.startswith("SYNCS") for o in ops)   # IQ ring: cp.async.bulk + mbarrier
        assert not any(i.startswith("LDG") and "iq" in i for i in body)
        assert any(".256" in i and i.startswith("STG") for i in body)
        assert any(i.startswith("BAR.ARV") for i in body) and any(i.startswith("BAR.SYNC") for i in body)
        assert sum(o.startswith("USETMAXREG") for o in ops) == 2
        assert ops.count("LDL") + ops.count("STL") == 0     # no spills at 88 (FFT) / 152 (interpolation) registers
    res = subprocess.run(["cuobjdump", "-res-usage", cw.lib_path()], capture_output=True, text=True, check=True).stdout
    regs = [int(m.group(1)) for m in re.finditer(r"demod_chan_kernel.*?\n.*?REG:(\d+)", res)]
    assert regs and max(regs) <= 120


def test_quantise_kernel_fits_beside_the_channelizer(cw):
    """The quantise pass runs in the 4096 registers per SM the channelizer leaves (128 threads x 32 registers): no
    spills, streaming 128-bit loads and stores, and an interior path that issues at least four loads before its first
    store (what keeps bytes in flight in that slot)."""
    funcs = _sass(cw)
    q = {k: v for k, v in funcs.items() if "quantise_kernel" in k}
    assert len(q) == 1
    body = next(iter(q.values()))
    ops = _ops(body)
    assert ops.count("LDL") + ops.count("STL") == 0
    wide = [i for i in body if (i.startswith("LDG") or i.startswith("STG")) and ".128" in i]
    first_store = next(n for n, i in enumerate(wide) if i.startswith("STG"))
    assert first_store >= 4 and all(".EF" in i for i in wide[:first_store])    # evict-first: read exactly once
    res = subprocess.run(["cuobjdump", "-res-usage", cw.lib_path()], capture_output=True, text=True, check=True).stdout
    regs = [int(m.group(1)) for m in re.finditer(r"quantise_kernel.*?\n.*?REG:(\d+)", res)]
    assert regs and max(regs) <= 32
