#!/usr/bin/env python
"""Benchmark of the hot path: Griffin-Lim vocoding of a batch of synthetic 80xT mels.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2|cfg5]

One "step" = one pass of the path over one batch: mel -> linear lift (pseudo-inverse, clamp,
^1.7) -> 60 Griffin-Lim iterations -> peak normalisation, for 32 utterances of 1000 frames
(BASELINE.json configs[1]; n_fft 1024, hop 256).  Metric: audio frames/s = sum(T) / time.

  value      whole-job frames/s with the mels already resident in HBM (device time, CUDA events
             around every step on the library's stream, max over ranks)
  e2e        the same from pinned HOST buffers through the public streaming call (xdtts_pipe_push, two
             batches in flight): every step's H2D of the mels and D2H of the waveforms are inside the timed
             region (the pipe is flushed before the clock stops) and overlap the neighbouring steps' kernels;
             e2e.blocking_call is the one-batch-at-a-time call (xdtts_gl_infer_batch), copies exposed
  roofline   the steady-state iteration kernel: algorithmic bytes (20K+8H per frame)
             / its mean launch duration inside the plan's CUDA graph (events around the graph of the full
             pass minus the graph of a half-length pass, per extra launch) vs the measured HBM peak of
             MEASURED_PEAKS.json; kernel_ms_stream_launches is the same kernel launched one by one
  cpu_baseline  the oracle's C/OpenMP port of the same step on this host's cores, bounded sample

--impl reference times that CPU port (the reference's Rust crate cannot be built here:
no cargo/rustc, crate not vendored -- DESIGN.md) as its own arm.
N > 1: launched by torchrun, one rank per GPU, 32 utterances per rank (weak scaling), no
data-path collective; ranks meet only at the barriers and the final all-reduce of the counters.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xd-tts_b200"))

CONFIGS = {
    # name: (batch per GPU, frames, n_fft, iterations)
    "cfg2": (32, 1000, 1024, 60),
    "cfg3": (32, 1000, 1024, 60),   # cfg2 preceded by the Tacotron2 postnet (BASELINE.json configs[2])
    "cfg5": (8, 8000, 2048, 60),
}
POSTNET_FLOP_PER_FRAME = 8683520.0   # 2 * 5 * (80*512 + 3*512*512 + 512*80)
FALLBACK_BF16_TFLOPS = 1590.0        # B200_PROFILING.md fallback (burst)
POWER, MOMENTUM, N_MELS, SR, FMAX = 1.7, 0.99, 80, 22050.0, 8000.0
STRONG_BATCH = 256                   # BASELINE.json configs[3]: batch = 256 sharded across the GPUs
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def workload_name(cfg, b, t, n_fft, it):
    pre = "Tacotron2 postnet (5 conv, bf16x3 tensor cores) + " if cfg == "cfg3" else ""
    return "%s: batch=%d x [80x%d] synthetic ln-mels U(-8,0), %spinv lift ^1.7 + Griffin-Lim %d iters, n_fft=%d hop=%d" % (
        cfg, b, t, pre, it, n_fft, n_fft // 4)


def measured_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if "bf16_tflops" in d:
                return float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
        except Exception:
            pass
    return FALLBACK_BF16_TFLOPS, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


def synth_postnet_layers(seed=7, channels=(80, 512, 512, 512, 512, 80), ksize=5):
    """Seeded random postnet (the LFS weights of the reference are absent): conv W ~ N(0, 1/sqrt(Cin k)),
    b ~ N(0, 0.1); BatchNorm gamma ~ U(.5, 1.5), beta, mean ~ N(0, 0.1), var ~ U(.5, 1.5)."""
    rng = np.random.default_rng(seed)
    layers = []
    for cin, cout in zip(channels[:-1], channels[1:]):
        layers.append(dict(
            w=(rng.standard_normal((cout, cin, ksize)) / np.sqrt(cin * ksize)).astype(np.float32),
            b=(0.1 * rng.standard_normal(cout)).astype(np.float32),
            gamma=rng.uniform(0.5, 1.5, cout).astype(np.float32),
            beta=(0.1 * rng.standard_normal(cout)).astype(np.float32),
            mean=(0.1 * rng.standard_normal(cout)).astype(np.float32),
            var=rng.uniform(0.5, 1.5, cout).astype(np.float32)))
    return layers


DEC_T_ENC, DEC_UNPADDED, DEC_STEPS = 100, 77, 1000   # the reference's fixed 100-position chunk, 1000-step cap (src/tacotron2/mod.rs:280,366)
# weights one decoder step reads: prenet 80x256 + 256x256, attention LSTM 4096x1792, query 128x1024, decoder LSTM
# 4096x2560, projection + gate 81x1536 (fp32)
DEC_WEIGHT_BYTES = 4 * (80 * 256 + 256 * 256 + 4096 * 1792 + 128 * 1024 + 4096 * 2560 + 81 * 1536)


def synth_decoder_weights(seed=11, gain=2.0):
    """Seeded random Tacotron2 decoder in PyTorch layouts (the LFS weights of the reference are absent)."""
    rng = np.random.default_rng(seed)

    def mat(rows, cols, g=gain):
        return (g * rng.standard_normal((rows, cols)) / np.sqrt(cols)).astype(np.float32)

    def vec(n, s=0.1):
        return (s * rng.standard_normal(n)).astype(np.float32)

    return dict(
        prenet1=mat(256, 80), prenet2=mat(256, 256), att_w_ih=mat(4096, 768), att_w_hh=mat(4096, 1024), att_b_ih=vec(4096),
        att_b_hh=vec(4096), query=mat(128, 1024), v=mat(1, 128, 2.0)[0],
        loc_conv=(rng.standard_normal((32, 2, 31)) / np.sqrt(62)).astype(np.float32), loc_dense=mat(128, 32, 1.0),
        dec_w_ih=mat(4096, 1536), dec_w_hh=mat(4096, 1024), dec_b_ih=vec(4096), dec_b_hh=vec(4096), proj_w=mat(80, 1536),
        proj_b=vec(80), gate_w=mat(1, 1536, 6.0), gate_b=np.array([-2.0], np.float32))


def synth_encoder_outputs(seed, t_enc):
    rng = np.random.default_rng(seed)
    return (0.5 * rng.standard_normal((t_enc, 512))).astype(np.float32), (0.5 * rng.standard_normal((t_enc, 128))).astype(np.float32)


def measured_l2_read_gbs(torch, device):
    """Read bandwidth of the L2 on this GPU, measured live: repeated reductions over a 64 MB tensor (fits the 126 MB L2,
    so every pass after the first is served from it).  A library kernel, used ONLY to calibrate the decoder's floor."""
    x = torch.empty(16 * 1024 * 1024, dtype=torch.float32, device="cuda:%d" % device).normal_()
    for _ in range(3):
        x.sum()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(5):
        e0.record()
        for _ in range(4):
            x.sum()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 4
        best = ms if best is None or ms < best else best
    return x.numel() * 4 / (best * 1e-3) / 1e9


DEC_HANDOVER_US = 0.55    # one hand-over between CTAs through the L2: a store's trip there + one poll round trip (~1050 cycles)
DEC_BARRIER_US = 1.23     # one split-phase grid barrier of this kernel with empty stages


def decoder_leg(ctx, with_cpu):
    """The decoder loop: 1000 steps of the Tacotron2 decoder in one persistent kernel, for one
    utterance (the reference's shape) and for 8 in lockstep.  Device time = CUDA events around the launch.
    Its 72.7 MB of weights never leave the chip (ncu: 0.15% DRAM throughput; a third of them stay in shared memory, the rest is
    re-read from the L2 every step) and a step is a chain of 7 dependent hand-overs between CTAs, so it is reported against
    (i) the hand-over floor and (ii) the bytes it reads from the L2 / the measured L2 read bandwidth -- not against HBM."""
    tacotron2 = ctx.tacotron2
    device = ctx.local_rank
    wt = synth_decoder_weights()
    dec = tacotron2.Decoder.from_weights(wt, gate_threshold=0.999999, max_steps=DEC_STEPS, seed=1, device=device)
    info = dec.info(1, DEC_T_ENC)
    out = {"bound": "latency: a step is a chain of %d dependent hand-overs between CTAs (polled {value, tag} cells through the L2 at batch "
                    "1-2, split-phase grid barriers at batch 4-8) and the serial stages between them; neither L2 nor HBM bandwidth" % info["handovers_per_step"],
           "kernel": "dec_persist_kernel (one cooperative launch = %d decoder steps)" % DEC_STEPS,
           "workload": "Tacotron2 decoder loop, t_enc=%d (%d unpadded), %d steps, fp32, synthetic weights" % (DEC_T_ENC, DEC_UNPADDED, DEC_STEPS),
           "weight_bytes_per_step": info["weight_bytes_per_step"], "weight_bytes_resident_in_smem": info["smem_resident_bytes"]}
    for nb in (1, 8):
        enc = [synth_encoder_outputs(100 + i, DEC_T_ENC) for i in range(nb)]
        best = None
        for _ in range(4):
            mels = dec.run_batch([m for m, _ in enc], [p for _, p in enc], [DEC_UNPADDED] * nb)
            ms, steps = dec.last_timing()
            best = ms if best is None or ms < best else best
        assert steps == DEC_STEPS and all(m.shape == (80, DEC_STEPS) and np.isfinite(m).all() for m in mels)
        out["b%d" % nb] = {"ms": best, "us_per_step": best * 1e3 / DEC_STEPS, "frames_per_s": nb * DEC_STEPS / (best * 1e-3),
                           "handover": "polled cells" if dec.info(nb, DEC_T_ENC)["polled_cells"] else "grid barriers"}
    l2 = measured_l2_read_gbs(ctx.torch, device)
    us = out["b1"]["us_per_step"]
    l2_bytes = info["weight_bytes_per_step"] - info["smem_resident_bytes"]
    l2_floor = l2_bytes / (l2 * 1e9) * 1e6
    ho_floor = info["handovers_per_step"] * DEC_HANDOVER_US
    out["floors"] = {"l2_read_gbs_measured": l2, "l2_bytes_per_step": l2_bytes, "l2_us_per_step": l2_floor, "handover_us_per_step": ho_floor,
                     "how": "L2: 64 MB torch reduction repeated in this process; hand-over: %d x %.2f us (a store's trip to the L2 + one poll "
                            "round trip; the grid barrier it replaced at batch 1-2 cost %.2f us with empty stages)" % (
                                info["handovers_per_step"], DEC_HANDOVER_US, DEC_BARRIER_US)}
    out["frac"] = l2_floor / us                     # the larger of the two floors: the bytes the step still reads from the L2
    out["frac_of_handover_floor"] = ho_floor / us
    out["achieved"] = l2_bytes / (us * 1e-6) / 1e9
    out["peak"] = l2
    out["unit"] = "GB/s (bytes read from the L2 per step / step time against the measured L2 read rate)"
    out["note"] = ("achieved / peak are L2 figures, NOT HBM: the weights never leave the chip (ncu: DRAM throughput 0.15%, L2 hit rate 93%).  "
                   "The step is latency-bound: with and without the shared-memory weight cache it differs by 1-2% (XDTTS_DEC_NO_CACHE=1).")
    if with_cpu:
        from oracle import decoder_oracle as d   # CPU baseline leg only

        mem, pm = synth_encoder_outputs(100, DEC_T_ENC)
        n = 40
        t0 = time.perf_counter()
        d.run_decoder(wt, mem, pm, DEC_UNPADDED, seed=1, gate_threshold=2.0, max_steps=n, dtype=np.float32)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "%d steps of the numpy/BLAS fp32 oracle (one utterance)" % n}
    dec.close()
    return out


def synth_mel(seed, n_mels, n_frames):
    """ln-mel in U(-8, 0), the range Tacotron2 emits: the benchmark's synthetic input."""
    return np.random.default_rng(seed).uniform(-8.0, 0.0, (n_mels, n_frames)).astype(np.float32)


def bench_mel(i, t):
    return synth_mel(1234 + i, N_MELS, t)    # utterance i of the job


def synth_batch(b, t, seed0):
    return [synth_mel(seed0 + i, N_MELS, t) for i in range(b)]


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for key in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
                if key in d:
                    return float(d[key]), "measured (MEASURED_PEAKS.json %s)" % key
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(cfg):
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(cfg)
        except Exception:
            return None
    return None


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, sampled in-process through NVML every
    5 ms (the nvidia-smi loop of B200_PROFILING.md cannot sample a region that lasts < 100 ms)."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.samples, self.mask, self.stop_flag, self.thread, self.h = index, [], 0, False, None, None
        self.sm_max = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def start(self):
        if self.h is None:
            return
        self.samples, self.mask, self.stop_flag = [], 0, False
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % getattr(self, "err", "?")]}
        self.stop_flag = True
        self.thread.join(timeout=1)
        reasons = sorted(name for name, bit in self.REASONS if self.mask & bit)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(self.samples), "how": "NVML, 5 ms period, timed region only"}


def bind_to_gpu_numa_node(index):
    """One process per GPU: run on the CPUs next to this GPU so that the pinned host buffers (first touch) sit on its
    NUMA node and the PCIe copies do not cross the socket interconnect.  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1 and 64 * w + b < n_cpu}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return sorted(allowed)
    except Exception:  # noqa: BLE001
        return None


def pinned_array(lib, shape):
    n = int(np.prod(shape))
    ptr = lib.xdtts_host_alloc(n * 4)
    if not ptr:
        raise MemoryError("xdtts_host_alloc")
    buf = (ctypes.c_float * n).from_address(ptr)
    return np.frombuffer(buf, dtype=np.float32).reshape(shape), ptr


# ------------------------------------------------------------------------------- CPU arm
def cpu_arm(cfg, steps, warmup, sample_utts=None, candidates=True):
    """The reference's path on this host's cores.  The Rust crate cannot be built here, so three CPU implementations of
    the same step are timed on a bounded sample and the FASTEST is quoted: the oracle's C/OpenMP
    port (rebuilt with -march=native on this host), torchaudio.functional.griffinlim (MKL FFT, all threads) and the
    numpy/scipy oracle (pocketfft, workers = cores).  Returns (frames/s, cores, kind, sample description, ms/step, all)."""
    from oracle import c_oracle, gl_oracle as o

    b, t, n_fft, it = CONFIGS[cfg]
    c_oracle.build()
    native = c_oracle.use_native_build()
    cores = c_oracle.num_threads()
    basis = o.create_mel_filter_bank(SR, n_fft, N_MELS, 0.0, FMAX)
    pinv = o.pinv_basis(basis).astype(np.float32)
    k = n_fft // 2 + 1
    # bounded sample of the workload: enough utterances to occupy every core, a few seconds per step
    n = sample_utts or max(2, min(b, cores))
    mels = np.stack(synth_batch(n, t, 1234))
    turns = np.stack([o.phase_turns(0, i, k, t) for i in range(n)])
    hop = n_fft // 4

    layers = synth_postnet_layers() if cfg == "cfg3" else None

    def cpu_postnet(m):   # numpy restatement of the ONNX graph (BLAS threads over the GEMMs), fp32
        from oracle.postnet_oracle import postnet

        return np.stack([postnet(x, layers, dtype=np.float32) for x in m]) if layers else m

    def per_utt():      # OpenMP over the frames of one utterance (rayon-style, as the crate does)
        pm = cpu_postnet(mels)
        for i in range(n):
            c_oracle.infer(pinv, pm[i], turns[i], hop, POWER, it, MOMENTUM)

    def per_batch():    # one thread per utterance
        c_oracle.infer_batch(pinv, cpu_postnet(mels), turns, hop, POWER, it, MOMENTUM)

    tried = {}
    best = None
    for name, fn in (("threads over utterances", per_batch), ("threads over frames", per_utt)):
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        tried["C/OpenMP port, " + name] = n * t / dt
        if best is None or dt < best[1]:
            best = (name, dt, fn)
    name, _, fn = best
    label = "C/OpenMP port of the oracle (%s), %s" % ("-march=native" if native else "x86-64-v3", name)
    if candidates:
        # the two library-backed implementations, on a smaller sample (they are several times slower)
        try:
            import torch
            import torchaudio

            torch.set_num_threads(cores)
            nb = min(n, 4)
            win = torch.hann_window(n_fft, periodic=True)

            def ta():
                pm = cpu_postnet(mels[:nb])
                s_lin = np.maximum(np.einsum("km,bmt->bkt", pinv, np.exp(pm)), 0.0) ** POWER
                torchaudio.functional.griffinlim(torch.from_numpy(s_lin.astype(np.float32)), win, n_fft, hop, n_fft, 1.0, it, MOMENTUM,
                                                 hop * (t - 1), True)

            ta()
            t0 = time.perf_counter()
            ta()
            tried["torchaudio.functional.griffinlim (%d threads)" % cores] = nb * t / (time.perf_counter() - t0)
        except Exception as e:  # noqa: BLE001
            tried["torchaudio.functional.griffinlim"] = "unavailable: %r" % (e,)
        try:
            def sp():
                pm = cpu_postnet(mels[:2])
                for i in range(2):
                    o.infer(pm[i], basis, n_fft - hop, POWER, it, MOMENTUM, turns[i], dtype=np.float32, workers=cores)

            t0 = time.perf_counter()
            sp()
            tried["numpy/scipy oracle (pocketfft, workers=%d)" % cores] = 2 * t / (time.perf_counter() - t0)
        except Exception as e:  # noqa: BLE001
            tried["numpy/scipy oracle"] = "unavailable: %r" % (e,)
    for _ in range(max(0, warmup - 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    sample = "%d of %d utterances x %d frames x %d iters per step, %s%s; fastest of the candidates" % (
        n, b, t, it, label, " after the numpy/BLAS postnet" if layers else "")
    return n * t / dt, cores, "port", sample, dt * 1e3, tried


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host thread it can
    if "TORCHELASTIC_RUN_ID" in os.environ or os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    cfg = args.config
    b, t, n_fft, it = CONFIGS[cfg]
    steps, warmup = args.steps, args.warmup
    v, cores, kind, sample, ms, tried = cpu_arm(cfg, steps, warmup)
    line = {
        "impl": "reference", "metric": "audio frames/sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, b, t, n_fft, it)},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample, "candidates": tried},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference Rust crate (griffin-lim 0.2.0) cannot be built here (no cargo, not vendored): CPU arm is the fastest of "
                "three CPU implementations of the same step (cpu_baseline.candidates, frames/s)",
    }
    emit(line)


# ------------------------------------------------------------------------------- GPU arm
class Ctx:
    pass


def measure(ctx, cfg, utt_ids, steps, warmup, *, kernel=True, blocking=True, pipe=True, tag=""):
    """One configuration on this rank's GPU: resident (device-timed) pass, the iteration kernel's steady-state duration,
    the lift, and the end-to-end calls.  Returns this rank's raw numbers (times in ms; the caller reduces over ranks)."""
    torch, lib, _ffi, griffin_lim, tacotron2, shard = ctx.torch, ctx.lib, ctx.ffi, ctx.griffin_lim, ctx.tacotron2, ctx.shard
    _, t, n_fft, it = CONFIGS[cfg]
    hop, k = n_fft // 4, n_fft // 2 + 1
    with_postnet = cfg == "cfg3"
    basis = griffin_lim.mel.create_mel_filter_bank(SR, n_fft, N_MELS, 0.0, FMAX)
    voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, POWER, it, MOMENTUM, seed=shard.rank_seed(0, ctx.rank), device=ctx.local_rank)
    mels = [bench_mel(i, t) for i in utt_ids]
    b = len(mels)
    frames = b * t
    r = {"b": b, "frames": frames, "t": t, "n_fft": n_fft, "it": it, "hop": hop, "k": k}
    plan = voc.plan([t] * b)
    r["info"] = plan.info()
    post = pplan = None
    if with_postnet:
        post = tacotron2.Postnet.from_layers(synth_postnet_layers(), device=ctx.local_rank)
        pplan = post.plan([t] * b)
        pplan.upload(mels)
    else:
        plan.upload(0, mels)

    def step():
        """one pass of the path over the resident batch -> device milliseconds"""
        ms_pn = pplan.run(feed=plan) if with_postnet else 0.0   # writes mel_outputs_postnet into the vocoder's arena
        return ms_pn + plan.run(0)[0]

    for _ in range(warmup):
        step()
    sampler = ClockSampler(ctx.local_rank)
    ctx.barrier()
    sampler.start()
    launches0 = lib.xdtts_kernel_launches()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(steps):
        dev_ms += step()
    ctx.barrier()
    r["wall_ms"] = (time.perf_counter() - t0) * 1e3
    r["dev_ms"] = dev_ms
    r["launches"] = lib.xdtts_kernel_launches() - launches0
    r["clocks"] = sampler.stop()

    if kernel:
        # steady-state duration of the iteration kernel, live, two ways:
        #  (1) as launched in the timed region -- inside the plan's CUDA graph: CUDA events around the graph of the full pass
        #      and around the graph of a pass with half the iterations (same batch, same kernels); the difference divided by
        #      the extra launches is the average duration of one steady-state launch where the product runs it;
        #  (2) kernel-by-kernel stream launches with events around the n_iter - 2 steady-state launches (adds the
        #      stream's launch-to-launch gap to every kernel: an upper bound, kept as kernel_ms_stream_launches).
        if with_postnet:
            plan.upload(0, mels)      # any mel will do for the timing: the arithmetic does not depend on the values
        iter_ms, iter_n, lift = 0.0, 0, []
        for _ in range(max(2, min(steps, 5))):
            _, mi, n = plan.run(_ffi.RUN_NO_GRAPH)
            iter_ms += mi
            iter_n += n
            lift.append(plan.lift_ms())
        r["kern_stream_ms"] = iter_ms / max(iter_n, 1)
        r["lift_single_ms"] = min(lift)                      # one launch bracketed by two events (+ launch / event latency)
        r["lift_ms"] = min(plan.time_lift(20) for _ in range(3))   # a train of 20 launches after a warm-up, per launch
        it_half = it // 2
        voc_half = griffin_lim.GriffinLim.new(basis, n_fft - hop, POWER, it_half, MOMENTUM, seed=shard.rank_seed(0, ctx.rank),
                                              device=ctx.local_rank)
        plan_half = voc_half.plan([t] * b)
        plan_half.upload(0, mels)
        for _ in range(2):
            plan_half.run(0)
            plan.run(0)
        t_full = min(plan.run(0)[0] for _ in range(5))
        t_half = min(plan_half.run(0)[0] for _ in range(5))
        r["kern_ms"] = (t_full - t_half) / (it - it_half)
        r["it_half"] = it_half
        plan_half.close()
        voc_half.close()

    out_len = hop * (t - 1)
    r["out_len"] = out_len
    pins = []
    if blocking or pipe:
        pin_in = [pinned_array(lib, (N_MELS, t)) for _ in range(b)]
        for (a, _), m in zip(pin_in, mels):
            a[...] = m
        pin_out = [pinned_array(lib, (out_len,)) for _ in range(b)]
        pins += pin_in + pin_out
        in_ptrs = _ffi.fptr_array([a for a, _ in pin_in])
        out_ptrs = _ffi.fptr_array([a for a, _ in pin_out])
        t_arr = (ctypes.c_int * b)(*([t] * b))
        peak_ok = True

    if blocking:
        # end to end through the blocking batch call, pinned and PAGEABLE host buffers (what a Rust caller's ndarray /
        # Vec buffers are: staged by the library's copy threads, overlapped with the DMA)
        def call(ip, op):
            if with_postnet:
                _ffi.check(lib.xdtts_tail_infer_batch(post._h, voc._h, ip, t_arr, b, None, None, op))
            else:
                _ffi.check(lib.xdtts_gl_infer_batch(voc._h, ip, t_arr, b, None, op))

        for _ in range(warmup):
            call(in_ptrs, out_ptrs)
        ctx.barrier()
        t1 = time.perf_counter()
        for _ in range(steps):
            call(in_ptrs, out_ptrs)
        ctx.barrier()
        r["call_ms"] = (time.perf_counter() - t1) * 1e3
        peak_ok = all(abs(float(np.abs(a).max()) - 1.0) < 1e-5 for a, _ in pin_out)
        pg_in = [np.array(m) for m in mels]
        pg_out = [np.zeros(out_len, np.float32) for _ in range(b)]      # touched once: the pages exist (a caller reusing its buffers)
        pg_ip, pg_op = _ffi.fptr_array(pg_in), _ffi.fptr_array(pg_out)
        for _ in range(warmup):
            call(pg_ip, pg_op)
        ctx.barrier()
        t1 = time.perf_counter()
        for _ in range(steps):
            call(pg_ip, pg_op)
        ctx.barrier()
        r["call_pageable_ms"] = (time.perf_counter() - t1) * 1e3
        peak_ok = peak_ok and all(abs(float(np.abs(a).max()) - 1.0) < 1e-5 for a in pg_out)

    if pipe:
        # end to end, streaming (the headline e2e): xdtts_pipe_push, two batches in flight, each with its own
        # pinned host buffers; every step's H2D and D2H are inside the timed region and overlap the kernels of
        # the neighbouring steps.  Same kernels, same results as the blocking call above (tests/test_gpu_postnet.py).
        pin_in2 = [pinned_array(lib, (N_MELS, t)) for _ in range(b)]
        for (a, _), m in zip(pin_in2, mels):
            a[...] = m
        pin_out2 = [pinned_array(lib, (out_len,)) for _ in range(b)]
        pins += pin_in2 + pin_out2
        sets = [(in_ptrs, out_ptrs), (_ffi.fptr_array([a for a, _ in pin_in2]), _ffi.fptr_array([a for a, _ in pin_out2]))]
        for a, _ in pin_out + pin_out2:
            a[...] = 0.0
        q = ctypes.c_void_p()
        _ffi.check(lib.xdtts_pipe_create(voc._h, post._h if with_postnet else None, t_arr, b, 2, ctypes.byref(q)))
        for i in range(warmup):
            _ffi.check(lib.xdtts_pipe_push(q, sets[i % 2][0], None, None, sets[i % 2][1]))
        _ffi.check(lib.xdtts_pipe_flush(q))
        ctx.barrier()
        t1 = time.perf_counter()
        for i in range(steps):
            _ffi.check(lib.xdtts_pipe_push(q, sets[i % 2][0], None, None, sets[i % 2][1]))
        _ffi.check(lib.xdtts_pipe_flush(q))
        ctx.barrier()
        r["e2e_ms"] = (time.perf_counter() - t1) * 1e3
        lib.xdtts_pipe_destroy(q)
        peak_ok = peak_ok and all(abs(float(np.abs(a).max()) - 1.0) < 1e-5 for a, _ in pin_out + pin_out2)
    if blocking or pipe:
        r["peak_ok"] = peak_ok

    if with_postnet:   # the tensor-core leg (BASELINE.json configs[2]): postnet alone, device time by CUDA events
        r["pn_ms"] = min(pplan.run(feed=plan) for _ in range(5))
        pplan.close()
        post.close()
    for _, ptr in pins:
        lib.xdtts_host_free(ptr)
    plan.close()
    voc.close()
    return r


def dropin_leg(ctx, with_cpu):
    """The reference's actual call: GriffinLim::infer(&mel) on ONE utterance with ordinary (pageable) host arrays
    (src/lib.rs:141; create_griffin_lim's 30 iterations, src/tacotron2/mod.rs:456) -- xdtts_gl_infer, B = 1."""
    griffin_lim, _ffi = ctx.griffin_lim, ctx.ffi
    n_fft, hop, it = 1024, 256, 30
    basis = griffin_lim.mel.create_mel_filter_bank(SR, n_fft, N_MELS, 0.0, FMAX)
    out = {"api": "xdtts_gl_infer (B = 1), pageable numpy buffers, %d iterations (the shipped configuration)" % it, "cases": {}}
    for t in (200, 1000):
        voc = griffin_lim.GriffinLim.new(basis, n_fft - hop, POWER, it, MOMENTUM, device=ctx.local_rank)
        mel = bench_mel(0, t)
        for _ in range(5):
            voc.infer(mel)
        n = 30
        t0 = time.perf_counter()
        for _ in range(n):
            y = voc.infer(mel)
        ms = (time.perf_counter() - t0) * 1e3 / n
        plan = voc.plan([t])
        plan.upload(0, [mel])
        for _ in range(3):
            plan.run(0)
        dev = min(plan.run(0)[0] for _ in range(10))
        # latency floor of this shape: 31 dependent launches, each at least one frame-chain long; the one-warp-per-run
        # mapping gives T / (resident warps) frames per warp, never fewer than 4
        info = plan.info()
        case = {"T": t, "ms_per_call": ms, "device_ms": dev, "frames_per_s": t / (ms * 1e-3),
                "realtime_factor": (ms * 1e-3) / (hop * (t - 1) / SR), "us_per_launch": dev * 1e3 / (it + 1),
                "runs": info["n_runs"], "run_frames": info["run_frames"], "peak_normalised_ok": bool(abs(float(np.abs(y).max()) - 1.0) < 1e-5)}
        if with_cpu:
            from oracle import c_oracle, gl_oracle as o

            c_oracle.build()
            pinv = o.pinv_basis(basis).astype(np.float32)
            tu = o.phase_turns(0, 0, n_fft // 2 + 1, t)
            c_oracle.infer(pinv, mel, tu, hop, POWER, it, MOMENTUM)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                c_oracle.infer(pinv, mel, tu, hop, POWER, it, MOMENTUM)
            cms = (time.perf_counter() - t0) * 1e3 / reps
            case["cpu_port_ms_per_call"] = cms
            case["cpu_port_cores"] = c_oracle.num_threads()
            case["speedup_vs_cpu_port"] = cms / ms
        out["cases"]["T%d" % t] = case
        plan.close()
        voc.close()
    return out


def run_gpu(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as g

    if not os.path.exists(g.LIB):
        g.build_cuda()
    from xdtts_b200 import _ffi, griffin_lim, shard, tacotron2

    ctx = Ctx()
    ctx.torch, ctx.lib, ctx.ffi, ctx.griffin_lim, ctx.tacotron2, ctx.shard = torch, _ffi.load_library(), _ffi, griffin_lim, tacotron2, shard
    ctx.rank, ctx.local_rank, ctx.world = rank, local_rank, world

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx.barrier = barrier
    cfg = args.config
    b, t, n_fft, it = CONFIGS[cfg]
    hop, k = n_fft // 4, n_fft // 2 + 1
    steps, warmup = args.steps, max(args.warmup, 3)
    strong = args.scaling == "strong"
    # the job: weak scaling = b utterances per GPU (world * b in total); strong = STRONG_BATCH utterances in total
    # (BASELINE.json configs[3]: batch = 256 sharded across the GPUs).  This rank vocodes its shard; no data crosses GPUs.
    total_utts = STRONG_BATCH if strong else world * b
    mine = shard.shard_utterances([t] * total_utts, world, rank)
    main = measure(ctx, cfg, mine, steps, warmup)

    def reduce(r, keys):
        vals = [r.get(key, 0.0) for key in keys]
        fr, la, vals = shard.reduce_counters(r["frames"], r.get("launches", 0), vals, device="cuda")   # the job's single collective
        return fr, la, dict(zip(keys, vals))

    keys = ["dev_ms", "wall_ms", "e2e_ms", "kern_ms", "call_ms", "call_pageable_ms", "kern_stream_ms", "lift_ms", "lift_single_ms"]
    total_frames, total_launches, red = reduce(main, keys)

    # ---- every N: the other scaling mode as a sub-object (resident + streaming e2e only), so that one run per N gives both
    # curves of BASELINE.json configs[3] ("weak and strong")
    other = None
    if cfg == "cfg2" and not args.no_extras:
        o_total = world * b if strong else STRONG_BATCH
        o_mine = shard.shard_utterances([t] * o_total, world, rank)
        o_r = measure(ctx, cfg, o_mine, max(3, steps // 2), 3, kernel=False, blocking=False, pipe=True)
        o_frames, _, o_red = reduce(o_r, ["dev_ms", "e2e_ms"])
        o_steps = max(3, steps // 2)
        other = {"scaling": "weak" if strong else "strong", "utterances_total": o_total, "utterances_this_gpu": len(o_mine),
                 "value": o_frames * o_steps / (o_red["dev_ms"] * 1e-3), "e2e": o_frames * o_steps / (o_red["e2e_ms"] * 1e-3),
                 "unit": "frames/s", "ms_per_step": o_red["dev_ms"] / o_steps, "runs": o_r["info"]["n_runs"],
                 "run_frames": o_r["info"]["run_frames"],
                 "note": "same kernels; the limiter of strong scaling is the per-GPU batch shrinking below one resident wave of "
                         "full-length runs (run_frames falls, run-boundary overhead rises), not communication: there is none"}

    # ---- N = 1 only: the other BASELINE configurations and the next-row legs, live in the same process
    extras = {}
    if world == 1 and cfg == "cfg2" and not args.no_extras:
        for c2, st in (("cfg3", max(3, steps // 2)), ("cfg5", max(3, steps // 4))):
            b2 = CONFIGS[c2][0]
            extras[c2] = (measure(ctx, c2, list(range(b2)), st, 3), st)
        extras["dropin"] = dropin_leg(ctx, not args.no_cpu)
    dec_line = decoder_leg(ctx, not args.no_cpu) if (world == 1 and cfg == "cfg2" and not args.no_decoder) else None

    if rank == 0:
        peak, peak_src = measured_peak()

        def roof(r, kern_ms):
            alg = r["frames"] * (20 * r["k"] + 8 * r["hop"])      # per steady-state launch on one GPU
            ach = alg / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
            return {"bound": "hbm", "kernel": "gl_iter_kernel<MID> (one Griffin-Lim iteration)", "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic(r["cfg"]),
                    "traffic_how": "static: dram__bytes_read + dram__bytes_write of one launch from the committed ncu --set full capture "
                                   "(profiles/roofline_traffic.json), not measured in this run",
                    "peak_source": peak_src, "kernel_ms": kern_ms,
                    "kernel_ms_how": "(CUDA-graph pass with %d iterations - pass with %d) / %d, CUDA events on the library's stream" % (
                        r["it"], r["it_half"], r["it"] - r["it_half"]),
                    "algorithmic_bytes_per_launch": alg}

        def lift_obj(r, lift_ms):
            alg = r["frames"] * 4 * (N_MELS + r["k"])            # 4 (80 + K) bytes per frame
            ach = alg / (lift_ms * 1e-3) / 1e9
            return {"bound": "hbm", "kernel": "gl_lift_tc_kernel (tcgen05 kind::f16, fp16 hi/lo split, 3 passes; exp prologue, clamp + ^1.7 epilogue)",
                    "ms": lift_ms, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                    "algorithmic_bytes": alg, "ms_single_bracketed_launch": r.get("lift_single_ms"),
                    "how": "CUDA events around a train of 20 launches on the library's stream after a warm-up, per launch (best of 3); "
                           "ms_single_bracketed_launch is one launch between two events in a kernel-by-kernel pass (best of 5) and "
                           "carries that pair's launch and event latency"}

        def postnet_obj(r):
            tpeak, tsrc = measured_tensor_peak()
            tf = POSTNET_FLOP_PER_FRAME * r["frames"] / (r["pn_ms"] * 1e-3) / 1e12
            return {"bound": "tensor", "kernel": "pn_conv_tc_kernel x5 + input staging (tcgen05 bf16x3: 3 MMA passes per layer)",
                    "ms": r["pn_ms"], "frames_per_s": r["frames"] / (r["pn_ms"] * 1e-3), "achieved": tf, "executed": 3 * tf,
                    "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak, "frac_executed": 3 * tf / tpeak,
                    "peak_source": tsrc, "algorithmic_flop_per_frame": POSTNET_FLOP_PER_FRAME, "in_timed_step": True}

        def e2e_obj(r, red_, st, frames_total):
            o_ = {"value": frames_total * st / (red_["e2e_ms"] * 1e-3), "unit": "frames/s",
                  "h2d_bytes_per_step": r["b"] * N_MELS * r["t"] * 4, "d2h_bytes_per_step": r["b"] * r["out_len"] * 4,
                  "ms_per_step": red_["e2e_ms"] / st,
                  "api": "xdtts_pipe_push%s, 2 batches in flight, pinned host buffers; flushed inside the timed region" % (
                      " (postnet + vocoder)" if r["cfg"] == "cfg3" else ""),
                  "peak_normalised_ok": r.get("peak_ok")}
            if "call_ms" in red_ and red_["call_ms"] > 0:
                name = "xdtts_tail_infer_batch" if r["cfg"] == "cfg3" else "xdtts_gl_infer_batch"
                o_["blocking_call"] = {"value": frames_total * st / (red_["call_ms"] * 1e-3), "ms_per_step": red_["call_ms"] / st,
                                       "api": name + " (one batch at a time, pinned host buffers)"}
                o_["blocking_call_pageable"] = {
                    "value": frames_total * st / (red_["call_pageable_ms"] * 1e-3), "ms_per_step": red_["call_pageable_ms"] / st,
                    "ratio_to_pinned": red_["call_ms"] / red_["call_pageable_ms"],
                    "api": name + " (one batch at a time, PAGEABLE host buffers -- a Rust caller's ndarray / Vec -- staged by the "
                                  "library's copy threads and overlapped with the DMA)"}
            return o_

        main["cfg"] = cfg
        main["it_half"] = main.get("it_half", it // 2)
        line = {
            "metric": "audio frames/sec", "value": total_frames * steps / (red["dev_ms"] * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": red["dev_ms"] / steps,
            "wall_ms_per_step": red["wall_ms"] / steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg, b, t, n_fft, it)},
            "plan": {"utterances_this_gpu": main["b"], "utterances_total": total_utts,
                     "l2": "no flush needed: per-step state %.0f MB > 126 MB L2" % ((20 * k + 8 * hop) * main["frames"] / 1e6),
                     "runs": main["info"]["n_runs"], "run_frames": main["info"]["run_frames"], "ctas": main["info"]["ctas"]},
            "e2e": e2e_obj(main, red, steps, total_frames),
            "gpu_launches": total_launches,
            # the same measurement in the other two units people quote for vocoders
            "frame_iters_per_s": total_frames * it * steps / (red["dev_ms"] * 1e-3),
            "realtime_factor": (red["dev_ms"] * 1e-3 / steps) / (total_utts * hop * (t - 1) / SR),
            "roofline": roof(main, red["kern_ms"]),
            "lift": lift_obj(main, red["lift_ms"]),
            "clocks": main["clocks"],
        }
        line["roofline"]["kernel_ms_stream_launches"] = red["kern_stream_ms"]
        if other is not None:
            line["strong_scaling" if other["scaling"] == "strong" else "weak_scaling"] = other
        if "pn_ms" in main:
            line["postnet"] = postnet_obj(main)
        for c2 in ("cfg3", "cfg5"):
            if c2 in extras:
                r2, st = extras[c2]
                r2["cfg"] = c2
                b2, t2, n2, it2 = CONFIGS[c2]
                red2 = {key: r2.get(key, 0.0) for key in keys}
                sub = {"config": {"workload": workload_name(c2, b2, t2, n2, it2)}, "value": r2["frames"] * st / (r2["dev_ms"] * 1e-3),
                       "unit": "frames/s", "steps": st, "ms_per_step": r2["dev_ms"] / st, "e2e": e2e_obj(r2, red2, st, r2["frames"]),
                       "roofline": roof(r2, r2["kern_ms"]), "lift": lift_obj(r2, r2["lift_ms"]), "clocks": r2["clocks"],
                       "plan": {"runs": r2["info"]["n_runs"], "run_frames": r2["info"]["run_frames"], "ctas": r2["info"]["ctas"]}}
                sub["roofline"]["kernel_ms_stream_launches"] = r2["kern_stream_ms"]
                if "pn_ms" in r2:
                    sub["postnet"] = postnet_obj(r2)
                line[c2] = sub
        if "dropin" in extras:
            line["dropin"] = extras["dropin"]
        if dec_line is not None:
            line["decoder"] = dec_line
        if world == 1 and not args.no_cpu:
            v, cores, kind, sample, _, tried = cpu_arm(cfg, 3, 1)   # ~15 s of CPU work incl. picking the fastest implementation
            line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample, "candidates": tried}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, torchrun
    notices), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the original."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-decoder", action="store_true", help="skip the decoder-loop leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg3 / cfg5 / drop-in / other-scaling sub-objects")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1: 32 utterances per GPU (weak) or 256 in total (strong)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
