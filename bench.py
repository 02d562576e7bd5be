#!/usr/bin/env python
"""bench.py -- CWSL_DIGI receive front-end on B200: channel-Msamples/s (IQ in x decoders).

Workload (BASELINE.json configs[4], the configuration the metric's per-GPU targets are quoted on):
  64 synthetic 192 kHz receivers x 1024 decoder channels each, one 15 s FT8 slot per step,
  demodFreq_c = -96000 + round(c*186000/1023); receivers are partitioned round-robin over the
  ranks (one process per GPU, no data-path collective) -> "scaling": "strong".
A step = every receiver of the rank: demodulation of all 1024 channels (NCO mix + 512-tap FIR + /16 + SSB demod),
max|x| -> normalise -> int16 for all channels (one launch), max reset (one launch).
--mode selects the demodulator behind `value`: stft (default; one FFT per output sample shared by all channels of
the receiver + per-channel interpolation, dynamic-range guard on, <= 1 LSB / >= 90 dB), fast (direct-form FFMA2
kernel, <= 1 LSB), exact (bit-identical to the reference).

One JSON line:
  value            : ch-samples/s with the IQ already resident in HBM (bind_device_iq), all ranks summed, --mode
  e2e              : same metric through the C ABI with HOST buffers: pinned IQ -> cwsl_rx_push_iq (H2D) ->
                     cwsl_rx_end_slot (kernels + D2H of the demodulated part of the int16 result); PCIe-bound
  roofline         : the dominant kernel of --mode (see roofline_by_mode)
  roofline_by_mode : every arithmetic mode timed on the same receivers with >= 5 steps each.
                     stft: demod_chan_kernel's algorithmic HBM bytes (IQ once + float audio once) against
                     MEASURED_PEAKS.json hbm_gbs, plus the shared-memory wavefront fraction (its real limiter);
                     fast / exact: FP32 FMA pipe (134 flop per channel-sample), peak = FFMA2
                     microbenchmark run in this process, issued-slot fraction next to the ncu counter
  configs          : BASELINE.json configs[0] (one FT8 decoder), [1] (default 20 m set), [3] (8 bands x {WSPR,
                     FST4W-1800}) host -> host through the C ABI; configs[2] (8 x 7 station) is `station`
  cpu_baseline / --impl reference: the reference's own SSBD.hpp/LowPass.hpp chain (oracle/_ref,
                     -O3 -mavx2 -mfma -ffast-math = the shipped /O2 /fp:fast /arch:AVX analogue), one worker thread
                     per host core, on a bounded SAMPLE of the workload (a rate: extrapolated to the full config).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS, IQ_LEN, PERIOD = 192000, 2048, 15.0
N_RECEIVERS, N_CHANNELS = 64, 1024
FLOP_PER_CH_SAMPLE = 134.0          # SURVEY.md section 8d (mix 6 + 32 taps x 4): the ALGORITHMIC work
# What demod_fast_kernel<16,4,128,2> actually issues on the FMA pipe per SSBD block (16 ch-samples), from its SASS
# (cuobjdump, fully unrolled 4-block row: 1028 FFMA2 + 280 FADD2 + 256 FMUL2, exchange adds included): the symmetric
# taps are folded (h[j] = h[512-j]) so fewer multiplies are executed than the algorithm counts. Each packed
# instruction = 2 lanes x 2 flop-slots; a segment of 3 tiles (1536 blocks) recomputes a 32-block overlap once.
FAST_PIPE_INSTR_PER_BLOCK = 391.0
FAST_TILE_OVERHEAD = 1536.0 / 1504.0
# demod_exact_tiled_kernel per SSBD block, every operation unfused in the reference's order (source/SSBD.hpp:160-183):
# mix 16 x (4 mul + 2 add) = 96; FIR rows 32 x 16 x (one packed mul = 2 lane-slots + 2 scalar adds) = 2048; row x
# phase 32 x 6 = 192; ascending-n gather 32 x 2 = 64  ->  2400 FMA-pipe lane-slots per block = 150 per channel-sample
EXACT_LANE_SLOTS_PER_CH_SAMPLE = 150.0
METRIC, UNIT = "channel-Msamples/s (IQ in x decoders)", "ch-Msamples/s"
NCU_COUNTERS = {   # ncu --set full captures committed under profiles/ (one launch each, 1024 ch x FT8 slot unless noted)
    "fast": dict(file="profiles/r2_demod_fast_ncu_full.csv", sm__pipe_fma_cycles_active_pct=82.8, issue_active_pct=48.4),
    "exact": dict(file="profiles/r2_demod_exact_tiled_ncu_full.csv", sm__pipe_fma_cycles_active_pct=77.8, issue_active_pct=64.0,
                  note="256-channel launch"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--receivers", type=int, default=N_RECEIVERS, help="total receivers over all ranks")
    ap.add_argument("--channels", type=int, default=N_CHANNELS)
    ap.add_argument("--mode", default="stft", choices=["stft", "fast", "exact"],
                    help="stft: FFT channelizer kernel + dynamic-range guard (<= 1 LSB, default); fast: direct-form FFMA2 "
                         "kernel (<= 1 LSB); exact: bit-identical to the reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-station", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE.json configs[0], [1], [3] legs")
    ap.add_argument("--no-other-modes", action="store_true",
                    help="skip the legs that time the other two arithmetic modes and check parity against EXACT")
    ap.add_argument("--per-receiver-streams", action="store_true",
                    help="resident arm: one CUDA stream per receiver (quantise of receiver r overlaps demod of r+1: "
                         "+2 %% throughput, but per-kernel event timings then overlap and the roofline figures are void)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-time budget of the cpu_baseline leg")
    return ap.parse_args()


def workload_name(a):
    return (f"stress sweep: {a.receivers} synthetic 192 kHz receivers x {a.channels} channels, one 15 s FT8 slot "
            f"per receiver per step (BASELINE.json configs[4])")


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "25", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.p = None

    def stop(self):
        if not self.p:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, reasons, pw = [], [], set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return dict(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                    reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the only places that execute oracle/)
# ------------------------------------------------------------------------------------------------
def host_iq(seed, n_blocks):
    """Same statistics as the GPU workload: Gaussian sigma=300 per component + a few strong carriers."""
    rng = np.random.default_rng(seed)
    n = n_blocks * IQ_LEN
    x = rng.standard_normal(2 * n, dtype=np.float32) * np.float32(300.0)
    t = np.arange(n, dtype=np.float64)
    for j in range(4):
        f = -90000 + 45000 * j + 1234
        ph = 2 * np.pi * ((f * t) % FS) / FS
        x[0::2] += (4000 * np.cos(ph)).astype(np.float32)
        x[1::2] += (4000 * np.sin(ph)).astype(np.float32)
    return x


def sample_note(n_ch, cores, flags, extra=""):
    return (f"SAMPLE, EXTRAPOLATED: 1 receiver x {n_ch} channels (every {N_CHANNELS // max(1, n_ch)}th of the stress set) x "
            f"one 15 s FT8 slot ({int(PERIOD * FS) // IQ_LEN * IQ_LEN} IQ samples) is timed and reported as a rate; the full "
            f"config (64 receivers x 1024 channels) would take {64 * 1024 // max(1, n_ch)} x as long on the same cores "
            f"(channels and receivers are independent, one thread per channel as source/Instance.cpp:173). Reference "
            f"SSBD.hpp/LowPass.hpp chain (oracle/_ref) built with {flags}, {cores} worker threads{extra}")


def cpu_chain(freqs, n_blocks, budget_s, fast=True, max_reps=50):
    """Time the reference chain (oracle/_ref) on all host cores. Returns dict for the JSON line."""
    from oracle.oracle import Ref, af_size
    kind = "reference"
    try:
        ref = Ref(fast=fast)
    except Exception as e:  # noqa: BLE001  prebuilt oracle/_ref missing: report it, never substitute anything
        return dict(value=None, unit=UNIT, cores=0, kind=kind, sample=f"oracle/_ref unavailable: {e}")
    cores = ref.hardware_concurrency() or os.cpu_count() or 1
    iq = host_iq(7, n_blocks)
    scales = np.full(len(freqs), 0.9, np.float32)
    afs = af_size(PERIOD)
    best, total, reps = None, 0.0, 0
    while reps < max_reps and (reps < 2 or total < budget_s):
        sec, _, _ = ref.chain_threads(FS, freqs, scales, iq, IQ_LEN, afs, max_threads=cores)
        total += sec
        reps += 1
        best = sec if best is None else min(best, sec)
        if sec > budget_s:
            break
    chs = float(len(freqs)) * n_blocks * IQ_LEN
    flags = "-O3 -mavx2 -mfma -ffast-math" if fast else "-O2 -ffp-contract=off (strict IEEE, parity build)"
    return dict(value=chs / best / 1e6, unit=UNIT, cores=int(cores), kind=kind, seconds=best, reps=reps, extrapolated=True,
                sample=sample_note(len(freqs), cores, flags, f", best of {reps}"))


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cwsl_digi_b200 import synth
    cores = os.cpu_count() or 1
    n_ch = int(min(a.channels, max(32, 4 * cores)))
    freqs = synth.stress_demod_freqs(a.channels)[:: max(1, a.channels // n_ch)][:n_ch].astype(np.int32)
    n_blocks = int(PERIOD * FS) // IQ_LEN
    secs = []
    from oracle.oracle import Ref, af_size
    ref = Ref(fast=True)
    cores = ref.hardware_concurrency() or cores
    iq = host_iq(7, n_blocks)
    scales = np.full(len(freqs), 0.9, np.float32)
    afs = af_size(PERIOD)
    chs = float(len(freqs)) * n_blocks * IQ_LEN
    for i in range(a.warmup + a.steps):
        sec, _, _ = ref.chain_threads(FS, freqs, scales, iq, IQ_LEN, afs, max_threads=cores)
        if i >= a.warmup:
            secs.append(sec)
    tot = sum(secs)
    value = chs * len(secs) / tot / 1e6
    sample = sample_note(len(freqs), cores, "-O3 -mavx2 -mfma -ffast-math", "; each step = one such sample")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=1e3 * tot / len(secs), higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=workload_name(a), sample=sample, extrapolated=True),
                cpu_baseline=dict(value=value, unit=UNIT, cores=int(cores), kind="reference", sample=sample, extrapolated=True),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                x_realtime_per_receiver=PERIOD / (tot / len(secs)) if len(freqs) else None, gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist

    import cwsl_digi_b200 as cw
    from cwsl_digi_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        a.gpus = world
    if not torch.cuda.is_available() or cw.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The communicator's first use prints "NCCL version ..." on stdout; stdout is reserved for the one JSON line,
        # so the file descriptor points at stderr while NCCL comes up.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    MODES = {"stft": cw.MODE_STFT, "fast": cw.MODE_FAST, "exact": cw.MODE_EXACT}
    mode = MODES[a.mode]
    from cwsl_digi_b200.sharding import receivers_of_rank
    my_rx = receivers_of_rank(a.receivers, rank, world)     # one receiver/band per GPU, round-robin
    n_blocks = int(PERIOD * FS) // IQ_LEN                   # 1406 IQ blocks = 2 879 488 samples
    n_iq = n_blocks * IQ_LEN
    freqs = synth.stress_demod_freqs(a.channels)
    afs = cw.af_size(PERIOD)
    n_sms = torch.cuda.get_device_properties(local).multi_processor_count
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (sustained)"
    except Exception:  # noqa: BLE001
        hbm_peak, hbm_src = 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md (MEASURED_PEAKS.json missing)"

    def prof_json(name):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:  # noqa: BLE001
            return {}

    # Synthetic IQ, one distinct array per receiver (seed 20261017 + receiver id), generated on the
    # device: complex white noise, sigma 300 per component, plus 8 complex tones of amplitude 8000 inside every
    # decoder passband at audio offsets drawn uniformly in [200, 2900] Hz (8192 tones at 1024 channels). The tones
    # are synthesised with one inverse FFT of slot length (their frequencies sit on its 0.067 Hz grid).
    rng = np.random.default_rng(synth.BASE_SEED)
    tone_hz = (freqs[:, None].astype(np.float64) + rng.uniform(200.0, 2900.0, (a.channels, 8))).reshape(-1)
    tone_bin = torch.from_numpy(np.mod(np.rint(tone_hz * n_iq / FS).astype(np.int64), n_iq)).cuda()
    iq_dev = []
    for r in my_rx:
        g = torch.Generator(device="cuda").manual_seed(synth.BASE_SEED + r)
        ph = torch.rand(tone_bin.numel(), device="cuda", generator=g, dtype=torch.float64) * (2 * np.pi)
        spec = torch.zeros(n_iq, device="cuda", dtype=torch.complex64)
        spec.index_add_(0, tone_bin, (8000.0 * torch.exp(1j * ph)).to(torch.complex64))
        z = torch.fft.ifft(spec, norm="forward")            # sum of the tones, no 1/n
        x = torch.randn(2 * n_iq, device="cuda", generator=g) * 300.0
        x[0::2] += z.real
        x[1::2] += z.imag
        iq_dev.append(x.contiguous())
        del spec, z, ph
    torch.cuda.synchronize()

    # The compute stream is a HIGH-priority stream: the receivers' post streams (quantise, D2H) are created with the
    # lowest priority by the library, so a demodulation that becomes runnable together with the previous receiver's
    # quantise pass gets its CTAs placed first and the quantise CTAs fill the rest of each SM.
    prio = 0 if os.environ.get("CWSL_STREAM_PRIORITIES") == "0" else -1
    try:
        stream = torch.cuda.Stream(priority=prio)
    except Exception:  # noqa: BLE001
        stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    free0 = torch.cuda.mem_get_info()[0]
    rxs = []
    for _ in my_rx:
        rx = cw.Receiver(local, FS, IQ_LEN, mode=mode)
        grp = rx.add_group(PERIOD)
        for f in freqs:
            rx.add_channel(grp, int(f), 0.9)
        if not a.per_receiver_streams:
            rx.set_stream(stream.cuda_stream)
        rxs.append(rx)
    # Default: all receivers of the rank are queued on ONE stream, so every demodulation kernel runs alone and its
    # CUDA-event duration is clean (roofline). --per-receiver-streams keeps each receiver's private stream instead
    # (what a station with one reader thread per receiver does); the timed region is bracketed by events on a master
    # stream that all receiver streams fork from / join into.
    rx_streams = [torch.cuda.ExternalStream(rx.stream()) for rx in rxs] if a.per_receiver_streams else []

    def step_resident():
        for rx, x in zip(rxs, iq_dev):
            rx.bind_device_iq(x.data_ptr(), n_blocks)
            rx.end_slot(0, None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fp32 = cw.measure_fp32_peak(local)
    peak_tf = max(fp32["ffma2_tflops"], fp32["ffma_tflops"])
    chs_step_total = float(a.receivers) * a.channels * n_iq

    def timed_resident(steps, warmup, sample_clocks):
        """W warm-up steps, then exactly `steps` steps between two events on the compute stream; the post streams'
        last quantise / max-reset passes are joined into the stream before the closing event."""
        for _ in range(warmup):
            step_resident()
        barrier()
        for rx in rxs:
            rx.enable_timing(True)
            rx.kernel_times()
        clocks = ClockSampler(local) if sample_clocks else None
        if clocks:
            clocks.start()
            time.sleep(0.3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s_ in rx_streams:                       # fork
            s_.wait_event(e0)
        for _ in range(steps):
            step_resident()
        for rx in rxs:                              # the last slots' post work belongs to the region
            rx.join_output()
        for s_ in rx_streams:                       # join
            ev = torch.cuda.Event()
            ev.record(s_)
            stream.wait_event(ev)
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        clk = clocks.stop() if clocks else None
        kt = [rx.kernel_times() for rx in rxs]
        for rx in rxs:
            rx.enable_timing(False)
        out = {k: sum(t[k] for t in kt) for k in kt[0]}
        out.update(ms_total=ms_total, ms_step=allmax(ms_total) / steps, clocks=clk, steps=steps)
        return out

    def isolated(n=4):
        """Each receiver alone on the device (outside any timed region): CUDA-event durations of the demodulator and
        the quantise pass that are not stretched by one waiting for the other's CTAs to drain."""
        acc = {}
        n = min(n, len(rxs))
        for rx, x in list(zip(rxs, iq_dev))[:n]:
            rx.synchronize()
            rx.enable_timing(True)
            rx.kernel_times()
            rx.bind_device_iq(x.data_ptr(), n_blocks)
            rx.end_slot(0, None)
            rx.synchronize()
            k = rx.kernel_times()
            rx.enable_timing(False)
            for key, v in k.items():
                acc[key] = acc.get(key, 0.0) + v
        return dict(demod_ms=acc["demod_ms"] / n, main_kernel_ms=acc["main_ms"] / n, guard_pre_ms=acc["guard_pre_ms"] / n,
                    guard_post_ms=acc["guard_post_ms"] / n, quantise_and_clear_ms=acc["quant_ms"] / n, receivers=n)

    def kernels_per_pass(name):
        # stft with its guard: band power + channelizer + selection + direct-form redo (indirect grid)
        return 4 if name == "stft" and a.channels >= 64 else 1

    def roofline_of(name, t, iso, sm_mhz):
        """Roofline block of one arithmetic mode from its own timed leg `t` (+ serialised pass `iso`)."""
        launches = max(1, int(t["demod_launches"]))
        ms_per_step = t["ms_step"]
        val = chs_step_total / (ms_per_step * 1e-3) / 1e6
        if name == "stft" and a.channels >= 64:
            launch_ms = t["main_ms"] / launches                 # demod_chan_kernel alone (events around it)
            stft_bytes = n_iq * 8 + a.channels * (n_iq // 16) * 4
            tj = prof_json("demod_chan_traffic.json")
            gbs = stft_bytes / (launch_ms * 1e-3) / 1e9
            wf = tj.get("lsu_shared_wavefronts_per_launch")
            cyc = launch_ms * 1e-3 * (sm_mhz or 1965.0) * 1e6 * n_sms
            cyc_iso = iso["main_kernel_ms"] * 1e-3 * (sm_mhz or 1965.0) * 1e6 * n_sms
            return dict(
                bound="hbm", achieved=gbs, peak=hbm_peak, unit="GB/s", frac=gbs / hbm_peak,
                traffic=tj.get("dram_bytes_per_launch"),
                kernel="demod_chan_kernel<16> (STFT channelizer: 8 FFT warps + 8 interpolation warps per SM, persistent)",
                limiter="no unit saturated (ncu, profiles/r2_demod_chan_final_ncu_full.csv): FMA pipe 54 %, shared-memory "
                        "wavefronts 57 %, issue slots 44 % with 16 resident warps per SM (registers + 215 KB of shared "
                        "memory cap the occupancy); the HBM fraction is what the contract asks for, the kernel is not HBM-bound",
                launch_ms=launch_ms, launches_timed=launches, value=val, ms_per_step=ms_per_step, steps=t["steps"],
                guard_ms_per_launch=(t["guard_pre_ms"] + t["guard_post_ms"]) / launches,
                kernel_share_of_step=iso["main_kernel_ms"] / (iso["demod_ms"] + iso["quantise_and_clear_ms"]),
                share_note="serialised share (each receiver alone on the device, CUDA events): comparable with the ncu "
                           "launch list under profiles/. In the timed region the quantise pass of receiver r runs on the "
                           "post stream beside the channelizer of r+1, and launch_ms there includes waiting for its CTAs",
                step_hbm=dict(
                    bytes_per_receiver=int(stft_bytes + a.channels * (n_iq // 16) * (4 + 2)),
                    achieved_gbs=(stft_bytes + a.channels * (n_iq // 16) * (4 + 2)) * len(rxs) / (ms_per_step * 1e-3) / 1e9,
                    frac=(stft_bytes + a.channels * (n_iq // 16) * (4 + 2)) * len(rxs) / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                    note="whole step: channelizer bytes + the normalise/quantise pass (float audio read, int16 written) of "
                         "every receiver of this rank / ms_per_step"),
                launch_ms_isolated=iso["main_kernel_ms"],
                achieved_isolated=stft_bytes / (iso["main_kernel_ms"] * 1e-3) / 1e9,
                frac_isolated=stft_bytes / (iso["main_kernel_ms"] * 1e-3) / 1e9 / hbm_peak,
                peak_source=hbm_src,
                algorithmic=f"{stft_bytes} B per launch = IQ {n_iq * 8} B read once + float audio "
                            f"{a.channels * (n_iq // 16) * 4} B written once (SURVEY.md 8d: 8/C + 4/16 B per channel-sample)",
                shared_memory=dict(
                    wavefronts_per_launch=wf, source=tj.get("source"),
                    frac_of_lsu_peak=(wf / cyc if wf else None), frac_of_lsu_peak_isolated=(wf / cyc_iso if wf else None),
                    ncu_pct_of_peak=tj.get("shared_wavefront_pct"), ncu_fma_pipe_pct=tj.get("fma_pipe_pct"),
                    ncu_issue_active_pct=tj.get("issue_active_pct"),
                    note="l1tex shared-memory wavefronts (ncu count of one launch, committed under profiles/) / (SMs x SM "
                         "cycles of launch_ms at the sampled clock), peak = 1 wavefront per SM per cycle"),
                direct_form_equivalent=dict(
                    tflops=FLOP_PER_CH_SAMPLE * val * 1e6 / 1e12, fp32_peak_tflops=peak_tf,
                    note="134 flop per channel-sample (SURVEY.md 8d, direct form) x throughput, for comparison with the "
                         "fast/exact modes only: the channelizer does not execute them"))
        launch_ms = t["demod_ms"] / launches
        achieved_tf = FLOP_PER_CH_SAMPLE * a.channels * n_iq / (launch_ms * 1e-3) / 1e12
        bytes_per_launch = n_iq * 8 + a.channels * (n_iq // 16) * (4 + 8)   # IQ once + float audio + phase table
        if name == "exact":
            slots = EXACT_LANE_SLOTS_PER_CH_SAMPLE
            kern = "demod_exact_tiled_kernel<16,128,3,24>"
            ex_note = ("FMA-pipe lane-slots of the reference's own operation list, every op unfused (packed mul = 2 slots, "
                       "scalar add = 1): 2400 per SSBD block; no tap folding is possible bit-exactly")
        else:
            slots = FAST_PIPE_INSTR_PER_BLOCK * 2 / 16 * FAST_TILE_OVERHEAD
            kern = "demod_fast_kernel<16,4,128,2>"
            ex_note = ("FMA-pipe lane-slots actually issued (391 packed f32x2 instructions per block from the SASS x 2 "
                       "lanes), tile overlap included")
        ex_tf = slots * 2 * a.channels * n_iq / (launch_ms * 1e-3) / 1e12
        return dict(
            bound="fp32_fma_pipe", achieved=achieved_tf, peak=peak_tf, unit="TFLOP/s", frac=achieved_tf / peak_tf,
            traffic=prof_json("demod_fast_traffic.json").get("dram_bytes_per_launch") if name == "fast" else None,
            kernel=kern, launch_ms=launch_ms, launches_timed=launches, value=val, ms_per_step=ms_per_step, steps=t["steps"],
            x_realtime_per_gpu=PERIOD * len(my_rx) / (ms_per_step * 1e-3),
            kernel_share_of_step=t["demod_ms"] / (t["ms_total"] if t["ms_total"] > 0 else 1),
            share_note="demod launch time / timed region, CUDA events; the normalise+quantise pass of receiver r runs on "
                       "its post stream under the demodulation of receiver r+1",
            peak_source="register-resident FMA microbenchmark with immediate operands run in this process "
                        "(cwsl_measure_fp32_peak, max of the FFMA2 and FFMA forms); MEASURED_PEAKS.json has no FP32-pipe "
                        f"figure. Nominal {n_sms} SM x 128 lanes x 2 x 1.965 GHz = {n_sms * 128 * 2 * 1.965e-3:.1f} TFLOP/s; "
                        f"measured FFMA2 {fp32['ffma2_tflops']:.1f}, scalar FFMA {fp32['ffma_tflops']:.1f}",
            algorithmic="134 flop per channel-sample (SURVEY.md 8d) x channels x IQ samples per launch"
                        + ("; frac can exceed 1: the kernel folds the symmetric taps and executes fewer multiplies than "
                           "the direct form the 134 counts (see 'executed')" if name == "fast" else ""),
            executed=dict(lane_slots_per_ch_sample=slots, tflops_equivalent=ex_tf, frac_of_peak=ex_tf / peak_tf, note=ex_note,
                          ncu=NCU_COUNTERS.get(name)),
            hbm=dict(achieved_gbs=bytes_per_launch / (launch_ms * 1e-3) / 1e9, peak_gbs=hbm_peak,
                     frac=bytes_per_launch / (launch_ms * 1e-3) / 1e9 / hbm_peak, bytes_per_launch=bytes_per_launch,
                     peak_source=hbm_src))

    # ---- device-resident arm, --mode ---------------------------------------------------------------
    main_t = timed_resident(a.steps, max(3, a.warmup), True)
    clk = main_t["clocks"]
    per_rx = a.channels * afs * (4 + 2)          # float audio scratch + int16 hand-off buffer of one receiver
    n_anchor = afs // 128 + 1                      # cwsl_kernels.hpp kChanAnchorHops
    footprint = dict(device_bytes_per_rank=int(free0 - torch.cuda.mem_get_info()[0]), receivers=len(rxs),
                     audio_and_handoff_bytes_per_receiver=int(per_rx),
                     phase_state_bytes=(dict(kind="anchors of the exact phase recurrence every 128 hops, shared by all "
                                                  "receivers with the same channel set (STFT mode keeps no full tables)",
                                             bytes=int(n_anchor * a.channels * 8))
                                        if a.mode == "stft" and a.channels >= 64 else
                                        dict(kind="full phase tables, 8 B per audio sample and distinct channel, shared by "
                                                  "all receivers", bytes=int((afs + 4) * a.channels * 8))),
                     note="cudaMemGetInfo delta over receiver creation + the timed steps of --mode (before the other "
                          "modes run): float audio scratch and int16 hand-off buffer of every receiver, phase state, "
                          "per-channel constants, STFT tables and guard scratch")
    iso_main = isolated() if rank == 0 else None
    barrier()
    ms_step = main_t["ms_step"]
    value = chs_step_total / (ms_step * 1e-3) / 1e6
    launches = int(kernels_per_pass(a.mode) * main_t["demod_launches"] + 2 * main_t["quant_launches"])

    # ---- end-to-end arm: host IQ in, host int16 out, through the C ABI ------------------------------
    e2e = None
    if not a.no_e2e:
        for rx in rxs:
            rx.synchronize()
        NBUF = min(int(os.environ.get("CWSL_BENCH_E2E_BUFFERS", "4")), len(rxs))
        host_in = [torch.empty(2 * n_iq, dtype=torch.float32).pin_memory() for _ in my_rx]
        for h, x in zip(host_in, iq_dev):
            h.copy_(x)
        # hand-off buffers from cwsl_host_alloc: pinned, and the library skips the known-zero tail on the wire
        host_out = [cw.HostBuffer(a.channels, afs) for _ in range(NBUF)]
        # Kernels of all receivers are queued FIFO on ONE compute stream (so receivers finish one after the other,
        # not all together), the H2D of the next receiver's IQ rides the same stream, and every finished slot is
        # copied back on its receiver's private copy stream (inside cwsl_rx_end_slot), overlapping the next kernels.
        try:
            e2e_stream = torch.cuda.Stream(priority=prio)
        except Exception:  # noqa: BLE001
            e2e_stream = torch.cuda.Stream()
        for rx in rxs:
            rx.set_stream(e2e_stream.cuda_stream)
        checks = [0]

        def step_e2e():
            for i, rx in enumerate(rxs):
                b = i % NBUF
                if i >= NBUF:
                    rxs[i - NBUF].wait_output()         # previous user of this pinned buffer has landed
                    checks[0] ^= int(host_out[b].array[0, 1000])   # consumer touches the result
                rx.push_iq((host_in[i].data_ptr(), n_blocks))
                rx.end_slot_packed(0, host_out[b].ptr)        # [channel][write_index]: one 1-D copy over PCIe
            for rx in rxs[-NBUF:]:
                rx.wait_output()
            e2e_stream.synchronize()

        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_e2e()
        torch.cuda.synchronize()
        wall = allmax(time.perf_counter() - t0)
        e2e = dict(value=chs_step_total * a.steps / wall / 1e6, unit=UNIT,
                   h2d_bytes_per_step=int(len(my_rx) * n_iq * 8 * world),
                   d2h_bytes_per_step=int(len(my_rx) * a.channels * (n_iq // 16) * 2 * world),
                   d2h_note="packed hand-off (cwsl_rx_end_slot_packed): int16 [channels][write_index = 179968] per receiver as one "
                            "contiguous copy; the zero tail of a decoder's 240000-sample buffer never crosses PCIe",
                   ms_per_step=1e3 * wall / a.steps,
                   d2h_gbs_per_gpu=len(my_rx) * a.channels * (n_iq // 16) * 2 * a.steps / wall / 1e9,
                   bound="PCIe: the int16 hand-off alone moves d2h_gbs_per_gpu over this GPU's Gen5 x16 link (1-D pinned copy "
                         "ceiling 55.6 GB/s, profiles/r2_d2h_probe.json; per rank count: profiles/r2_numa_probe_n*.json); the "
                         "kernels need 1/8 of that time",
                   note="pinned host IQ -> cwsl_rx_push_iq -> cwsl_rx_end_slot_packed(host int16); timed region includes "
                        "every H2D and D2H copy; wall clock around a device synchronize, max over ranks; "
                        f"{NBUF} pinned hand-off buffers in rotation")
        for h in host_out:
            h.free()
        del host_in
        for rx in rxs:
            rx.synchronize()
            rx.set_stream(stream.cuda_stream)

    # ---- every arithmetic mode on the same receivers (>= 5 timed steps each) + in-run parity against EXACT ----
    roofline_by_mode, parity = {}, None
    sm_mhz = (clk or {}).get("sm_mhz")
    if rank == 0:
        roofline_by_mode[a.mode] = roofline_of(a.mode, main_t, iso_main, sm_mhz)
    if not a.no_other_modes and not a.per_receiver_streams:
        for name, m in MODES.items():
            if name == a.mode:
                continue
            for rx in rxs:
                rx.synchronize()
                rx.set_mode(m)
            t = timed_resident(max(5, a.steps), 2, False)
            iso = isolated(2) if rank == 0 else None
            barrier()
            if rank == 0:
                roofline_by_mode[name] = roofline_of(name, t, iso, sm_mhz)
        if rank == 0:
            # receiver 0 of this rank, every channel, on the bench's own IQ: int16 of each mode against the EXACT
            # mode's (which the GPU tests pin bit-for-bit to the reference chain); float residual on 32 channels.
            # Second input: one carrier 80 dB above the receiver noise, every other channel quiet (the STFT guard's case).
            rx0, sel = rxs[0], list(range(0, a.channels, max(1, a.channels // 32)))
            wi = n_blocks * IQ_LEN // 16

            def compare(x_dev, label):
                got = {}
                for name, m in MODES.items():
                    rx0.set_mode(m)
                    rx0.bind_device_iq(x_dev.data_ptr(), n_blocks)
                    rx0.end_slot(0, None)
                    rx0.synchronize()
                    q = torch.empty((a.channels, afs), dtype=torch.int16, device="cuda")
                    for c in range(a.channels):
                        rx0.copy_device_audio(0, c, q[c].data_ptr())
                    rx0.synchronize()
                    got[name] = (q, {c: rx0.read_float_audio(0, c) for c in sel},
                                 rx0.guard_stats(0) if name == "stft" and a.channels >= 64 else None)
                out = {"input": label}
                band_rms = float(torch.sqrt(torch.mean(x_dev.double() ** 2) * 2).item())   # rms of |x| over the slot
                ex_rms = {c: np.sqrt(np.mean(got["exact"][1][c][:wi].astype(np.float64) ** 2)) for c in sel}
                loud = max(sel, key=lambda c: ex_rms[c])
                for name in ("fast", "stft"):
                    d = (got[name][0].to(torch.int32) - got["exact"][0].to(torch.int32)).abs()
                    worst, worst_band, loud_db = -1e9, -1e9, None
                    for c in sel:
                        want = got["exact"][1][c][:wi].astype(np.float64)
                        err = got[name][1][c][:wi].astype(np.float64) - want
                        e_rms = max(np.sqrt(np.mean(err ** 2)), 1e-300)
                        r = 20 * np.log10(e_rms / max(ex_rms[c], 1e-300))
                        worst = max(worst, r)
                        worst_band = max(worst_band, 20 * np.log10(e_rms / band_rms))
                        if c == loud:
                            loud_db = r
                    out[name] = dict(max_int16_lsb=int(d.max().item()), differing_samples=float((d > 0).float().mean().item()),
                                     worst_residual_db=float(worst), residual_db_strongest_channel=float(loud_db),
                                     worst_error_vs_band_rms_db=float(worst_band))
                    if got[name][2]:
                        out[name]["guard"] = got[name][2]
                quiet = min(sel, key=lambda c: ex_rms[c])
                out["levels"] = dict(band_rms=band_rms, strongest_channel_audio_rms=float(ex_rms[loud]),
                                     quietest_channel_audio_rms=float(ex_rms[quiet]),
                                     quietest_channel_db_below_band=float(20 * np.log10(max(ex_rms[quiet], 1e-300) / band_rms)))
                return out

            parity = {"reference": "EXACT mode (bit-identical to the reference chain, tests/test_parity_gpu.py)",
                      "channels": "all channels (int16) / every 32nd channel (float residual)",
                      "bench_input": compare(iq_dev[0], "receiver 0 of this run")}
            g = torch.Generator(device="cuda").manual_seed(80)
            xh = torch.randn(2 * n_iq, device="cuda", generator=g) * 3.0
            tt = torch.arange(n_iq, device="cuda", dtype=torch.float64)
            ph = 2 * np.pi * ((int(freqs[a.channels // 3]) + 1500) * tt % FS) / FS
            xh[0::2] += (3.0e4 * torch.cos(ph)).float()
            xh[1::2] += (3.0e4 * torch.sin(ph)).float()
            parity["high_dynamic_range"] = compare(xh.contiguous(), "noise sigma 3 + one carrier of amplitude 30000 "
                                                   "(+80 dB) in one passband, all other channels quiet")
            parity["high_dynamic_range"]["note"] = (
                "worst_residual_db is relative to each channel's OWN audio. A quiet channel here holds the stop-band leakage "
                "of a carrier ~80 dB above it, so float32 rounding of that carrier's partial sums (~1e-7 of the band, see "
                "worst_error_vs_band_rms_db) is all that separates two correct float32 evaluations of the reference's "
                "formula in different summation order: FAST against EXACT shows the same figure as STFT, whose guard has "
                "handed these channel segments to the FAST kernel (guard.redone). The <= 1 int16 LSB bar holds on every "
                "channel; the >= 90 dB bar holds wherever the channel is within ~60 dB of the band's strongest signal "
                "(residual_db_strongest_channel, and the bench input above)")
            del xh, tt, ph
        for rx in rxs:
            rx.set_mode(mode)
            rx.synchronize()
    barrier()

    # ---- BASELINE.json configs[0], [1], [3]; configs[2] = the 8-receiver x 7-mode skimmer station, streamed ----
    for rx in rxs:
        rx.close()
    rxs = []
    del iq_dev
    torch.cuda.empty_cache()
    configs = None
    if not a.no_configs:
        configs = run_small_configs(cw, torch, rank, world, local, mode, barrier, allmax)
    station = None
    if not a.no_station:
        station = run_station(cw, torch, dist, rank, world, local, mode)

    # ---- station-level gather (NCCL): rank 0 collects one channel's audio per rank ------------------------
    gathered = None
    if world > 1:
        with cw.Receiver(local, FS, IQ_LEN, mode=mode) as rxg:
            gg = rxg.add_group(PERIOD)
            rxg.add_channel(gg, -26000 + 100 * rank, 0.9)
            xg = (torch.randn(2 * 64 * IQ_LEN, device="cuda") * 300.0).contiguous()
            rxg.bind_device_iq(xg.data_ptr(), 64)
            rxg.end_slot(gg, None)
            sample = torch.empty(afs, dtype=torch.int16, device="cuda")
            rxg.copy_device_audio(gg, 0, sample.data_ptr())
            rxg.synchronize()
            from cwsl_digi_b200.sharding import gather_slot_audio
            bucket = gather_slot_audio(sample, dst=0)
            if rank == 0:
                gathered = [int(b.to(torch.int64).abs().sum().item()) for b in bucket]

    if rank == 0:
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n_ch = int(min(a.channels, max(32, 4 * cores)))
            sub = freqs[:: max(1, a.channels // n_ch)][:n_ch].astype(np.int32)
            cpu = cpu_chain(sub, n_blocks, a.cpu_seconds, fast=True)
            strict = cpu_chain(sub[: max(8, cores)], n_blocks, a.cpu_seconds / 2, fast=False, max_reps=3)
            cpu["strict_build"] = dict(value=strict["value"], seconds=strict.get("seconds"), sample=strict["sample"])
        other_modes = {k: dict(value=v["value"], unit=UNIT, ms_per_step=v["ms_per_step"], steps=v["steps"])
                       for k, v in roofline_by_mode.items() if k != a.mode}
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=max(3, a.warmup),
                    ms_per_step=ms_step, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                    data="synthetic",
                    config=dict(workload=workload_name(a), receivers_total=a.receivers, channels=a.channels,
                                receivers_per_rank=len(my_rx), sample_rate=FS, iq_len=IQ_LEN, slot_s=PERIOD,
                                input="SURVEY.md 8d: complex noise sigma 300 + 8 tones of amplitude 8000 in every decoder passband",
                                mode=a.mode, parallelism=f"receiver-sharded x{world}, no data-path collective",
                                streams="one CUDA stream per receiver, forked from / joined into the timed master stream"
                                        if a.per_receiver_streams else "one stream for all receivers of the rank; "
                                        "post streams (quantise, max reset) joined before the closing event",
                                l2="inputs larger than L2: every launch reads a different receiver's 23 MB IQ and "
                                   "writes 737 MB of float audio; a step touches > 60 GB"),
                    x_realtime_per_gpu=PERIOD * len(my_rx) / (ms_step * 1e-3),
                    gchs_per_gpu=value / 1e3 / world,
                    clocks=clk, e2e=e2e, gpu_launches=launches, roofline=roofline_by_mode.get(a.mode),
                    roofline_by_mode=roofline_by_mode, cpu_baseline=cpu,
                    kernel_ms=dict(demod=main_t["demod_ms"], main_kernel=main_t["main_ms"], guard_pre=main_t["guard_pre_ms"],
                                   guard_post=main_t["guard_post_ms"], quantise_and_clear=main_t["quant_ms"],
                                   event_total=main_t["ms_total"], isolated_per_receiver=iso_main),
                    device_footprint=footprint, configs=configs, station=station, other_modes=other_modes,
                    parity_in_run=parity, gathered_checksums=gathered)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


STATION_MODES = [("FT8", 15.0, 0.90), ("FT4", 7.5, 0.90), ("JT65", 60.0, 0.90), ("WSPR", 120.0, 0.20),
                 ("FST4-120", 120.0, 0.90), ("FST4W-120", 120.0, 0.90), ("JS8", 15.0, 0.90)]
LO_20M = 14100000
CONFIG1_DECODERS = [(14095600, 120.0, 0.2), (14090000, 15.0, 0.9), (14080000, 7.5, 0.9), (14074000, 15.0, 0.9),
                    (14076000, 60.0, 0.9), (14078000, 15.0, 0.9), (14097000, 120.0, 0.9)]   # SURVEY.md 8d config 2 list


def run_small_configs(cw, torch, rank, world, local, mode, barrier, allmax):
    """BASELINE.json configs[0], [1] on rank 0 and configs[3] sharded by receiver, all host -> host through the C ABI
    (pinned host IQ, cwsl_rx_push_iq, cwsl_rx_end_slot into managed host buffers), wall clock, best of the timed passes.
    Small channel groups run the direct-form kernel in stft mode as well (break-even at 64 channels)."""
    from cwsl_digi_b200.sharding import receivers_of_rank
    mode_name = {cw.MODE_EXACT: "exact", cw.MODE_FAST: "fast", cw.MODE_STFT: "stft (groups < 64 channels: FAST kernel)"}[mode]
    out = {}

    def host_to_host(decs, seconds, chunk_s, m, reps=5):
        nblk = int(seconds * FS) // IQ_LEN
        g = torch.Generator(device="cuda").manual_seed(4242)
        x = torch.randn(nblk * IQ_LEN * 2, device="cuda", generator=g) * 300.0
        h = torch.empty(nblk * IQ_LEN * 2, dtype=torch.float32).pin_memory()
        h.copy_(x)
        del x
        rx = cw.Receiver(local, FS, IQ_LEN, mode=m)
        groups = {}
        for dial, per, sc in decs:
            if per not in groups:
                groups[per] = rx.add_group(per)
            rx.add_channel(groups[per], dial - LO_20M, sc)
        outs = {per: cw.HostBuffer(rx.num_channels(g_), rx.group_af_size(g_)) for per, g_ in groups.items()}
        step = int(chunk_s * FS) // IQ_LEN
        times = []
        for it in range(reps + 2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pos, k = 0, 0
            while pos < nblk:
                nb = min(step, nblk - pos)
                rx.push_iq((h.data_ptr() + pos * IQ_LEN * 8, nb))
                pos += nb
                k += 1
                for per, g_ in groups.items():
                    if abs((k * chunk_s) / per - round((k * chunk_s) / per)) < 1e-9:
                        rx.end_slot(g_, outs[per].ptr)
            rx.synchronize()
            if it >= 2:
                times.append(time.perf_counter() - t0)
        rx.close()
        for b in outs.values():
            b.free()
        best = min(times)
        return dict(seconds_of_signal=seconds, decoders=len(decs), wall_ms=1e3 * best, wall_ms_median=1e3 * statistics.median(times),
                    x_realtime=seconds / best, value=len(decs) * nblk * IQ_LEN / best / 1e6, unit=UNIT, passes=reps)

    if rank == 0:
        for key, decs, secs, label in (
                ("config0", [(14074000, 15.0, 0.9)], 15.0,
                 "BASELINE.json configs[0]: one 192 kHz receiver, single FT8 decoder at 14074000, one 15 s slot"),
                ("config1", CONFIG1_DECODERS, 120.0,
                 "BASELINE.json configs[1]: one 192 kHz receiver on 20 m, default decoder set (WSPR, FT8 x2, FT4, JT65, "
                 "JS8, FST4W-120), 120 s of signal, slot edges per mode")):
            r = host_to_host(decs, secs, 7.5, mode)
            r["exact"] = {k: v for k, v in host_to_host(decs, secs, 7.5, cw.MODE_EXACT, reps=3).items()
                          if k in ("wall_ms", "x_realtime", "value")}
            r.update(workload=label, mode=mode_name,
                     note="pinned host IQ -> host int16, every copy inside the timed region; H2D-bound")
            out[key] = r
    barrier()

    # configs[3]: 8 bands x {WSPR 120 s, FST4W-1800}; one 1800 s hyper-period per receiver = 345.6 M IQ samples
    # (2.76 GB) streamed from a 120 s pinned host buffer that is pushed 15 times; WSPR slots end every 120 s, the
    # FST4W-1800 slot once. Phase tables: 1.44 M entries (WSPR) and 21.6 M entries = 173 MB (FST4W-1800), shared by
    # the receivers (same demodulation frequencies on every band).
    mine = receivers_of_rank(8, rank, world)
    free0 = torch.cuda.mem_get_info()[0]
    chunk_blocks = int(120.0 * FS) // IQ_LEN
    rxs, hosts, outs = [], [], []
    for r in mine:
        rx = cw.Receiver(local, FS, IQ_LEN, ring_seconds=8.0, mode=mode)
        g_w = rx.add_group(120.0)
        rx.add_channel(g_w, -4400, 0.2)            # WSPR
        g_f = rx.add_group(1800.0)
        rx.add_channel(g_f, -3000, 0.9)            # FST4W-1800
        g = torch.Generator(device="cuda").manual_seed(999 + r)
        x = torch.randn(2 * chunk_blocks * IQ_LEN, device="cuda", generator=g) * 300.0
        h = torch.empty(2 * chunk_blocks * IQ_LEN, dtype=torch.float32).pin_memory()
        h.copy_(x)
        del x
        rxs.append((rx, g_w, g_f))
        hosts.append(h)
        outs.append((cw.HostBuffer(1, rx.group_af_size(g_w)), cw.HostBuffer(1, rx.group_af_size(g_f))))

    def long_pass():
        for k in range(15):
            for (rx, g_w, g_f), h, (ow, of) in zip(rxs, hosts, outs):
                rx.push_iq((h.data_ptr(), chunk_blocks))
                rx.end_slot(g_w, ow.ptr)
                if k == 14:
                    rx.end_slot(g_f, of.ptr)
        for rx, _, _ in rxs:
            rx.synchronize()

    long_pass()                                     # warm-up: tables, ring, first launches
    foot = int(free0 - torch.cuda.mem_get_info()[0])
    times = []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        long_pass()
        times.append(allmax(time.perf_counter() - t0))
    checks = [int(np.abs(of.array[0, :1000].astype(np.int64)).sum()) for _, of in outs]
    for rx, _, _ in rxs:
        rx.close()
    for ow, of in outs:
        ow.free()
        of.free()
    del hosts
    best = min(times)
    iq_per_rx = 15 * chunk_blocks * IQ_LEN
    if rank == 0:
        out["config3"] = dict(
            workload="BASELINE.json configs[3]: 8 bands x {WSPR 120 s, FST4W-1800}, 1800 s of signal per receiver "
                     "(345.6 M IQ samples = 2.76 GB each) streamed from pinned host memory in 120 s pushes through an 8 s "
                     "device ring; 15 WSPR slots + 1 FST4W-1800 slot per receiver copied back to the host",
            mode=mode_name, wall_s=best, passes=len(times), x_realtime=1800.0 * 1.0 / best, receivers_per_rank=len(mine),
            value=8 * 2 * iq_per_rx / best / 1e6, unit=UNIT, h2d_bytes=8 * iq_per_rx * 8,
            h2d_gbs_per_gpu=len(mine) * iq_per_rx * 8 / best / 1e9,
            device_footprint_bytes_per_rank=foot,
            footprint_note="cudaMemGetInfo delta: per receiver the 8 s IQ ring (12.3 MB), float audio 6.0 + 86.6 MB, int16 "
                           "3.0 + 43.3 MB; shared by all receivers of the rank: phase tables 11.5 MB (WSPR) + 173 MB (FST4W-1800)",
            result_checksums=checks)
    torch.cuda.empty_cache()
    return out if rank == 0 else None


def run_station(cw, torch, dist, rank, world, local, mode, n_receivers=8, hyper_s=120.0, chunk_s=1.5):
    """8 receivers (bands) x 7 mode families = 56 decoders (README.md:3), receiver r on rank r mod world.
    One 120 s hyper-period of signal per receiver is streamed from pinned HOST memory in 1.5 s pushes through
    a 3 s device ring; every mode's slot ends on its own period boundary (8 FT8, 16 FT4, 2 JT65, 1 WSPR ...),
    the int16 audio of each finished slot is copied back to the host. Reports x real time."""
    from cwsl_digi_b200.sharding import receivers_of_rank
    mine = receivers_of_rank(n_receivers, rank, world)
    blocks_per_chunk = int(chunk_s * FS) // IQ_LEN            # 140 IQ blocks = 1.493 s
    n_chunks = int(hyper_s / chunk_s)
    n_blocks = blocks_per_chunk * n_chunks
    rxs, hosts, outs, groups = [], [], [], []
    for r in mine:
        rx = cw.Receiver(local, FS, IQ_LEN, ring_seconds=3.0, mode=mode)
        gmap = {}
        for j, (name, per, sc) in enumerate(STATION_MODES):
            if per not in gmap:
                gmap[per] = rx.add_group(per)
            rx.add_channel(gmap[per], -80000 + 20000 * j + 1000 * r, sc)
        rxs.append(rx)
        groups.append(gmap)
        g = torch.Generator(device="cuda").manual_seed(777 + r)
        x = torch.randn(2 * n_blocks * IQ_LEN, device="cuda", generator=g) * 300.0
        h = torch.empty(2 * n_blocks * IQ_LEN, dtype=torch.float32).pin_memory()
        h.copy_(x)
        hosts.append(h)
        outs.append({per: cw.HostBuffer(rx.num_channels(gi), rx.group_af_size(gi)) for per, gi in gmap.items()})
        del x
    torch.cuda.synchronize()

    def one_pass():
        slots = 0
        for ck in range(n_chunks):
            t_end = (ck + 1) * chunk_s
            for rx, h, gmap, ob in zip(rxs, hosts, groups, outs):
                rx.push_iq((h.data_ptr() + ck * blocks_per_chunk * IQ_LEN * 8, blocks_per_chunk))
                for per, gi in gmap.items():
                    if abs((t_end / per) - round(t_end / per)) < 1e-9:      # this mode's slot edge
                        rx.end_slot(gi, ob[per].ptr)
                        slots += 1
        for rx in rxs:
            rx.synchronize()
        return slots

    one_pass()                                                  # warm-up (allocations, first launches)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    slots = one_pass()
    wall = time.perf_counter() - t0
    tw = torch.tensor([wall], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    wall = float(tw.item())
    for rx in rxs:
        rx.close()
    for ob in outs:
        for b in ob.values():
            b.free()
    chs = float(n_receivers) * len(STATION_MODES) * n_blocks * IQ_LEN
    return dict(workload="8 receivers x 7 modes = 56 decoders (BASELINE.json configs[2]), 120 s of signal per receiver "
                         "streamed from pinned host memory in 1.5 s pushes through a 3 s device ring, slot edges per mode, "
                         "int16 audio of every slot copied back to the host",
                x_realtime=hyper_s / wall, wall_s=wall, value=chs / wall / 1e6, unit=UNIT,
                receivers_per_rank=len(mine), slots_finished_per_rank=slots, mode={0: "exact", 1: "fast", 2: "stft (7-channel groups: FAST kernel)"}[mode])


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
