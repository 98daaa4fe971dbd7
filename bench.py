#!/usr/bin/env python
"""bench.py - mixture audio-seconds per second through STFT -> mask -> iSTFT.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one pass of the hot path over one batch of synthetic mixtures
(BASELINE.json config C2 per GPU: 256 x 3 s @ 16 kHz, FFT 512, hop 128, S = 3):

    log-feature = stft_log(wave)            (kernel 1: A1+A2+A3)
    [separator stand-in: per-source masks already resident in HBM]
    waves       = mask_istft(wave, mask)    (kernel 2: A1+A7+A8)

`value` is device-timed (CUDA events, max over ranks) with inputs resident in HBM;
`e2e` is the same step through the host-buffer API (pinned host waves in, pinned
host waveforms out, copies inside the timed region).  N > 1: one process per GPU
(torchrun), every rank runs its own C2 batch (weak scaling, no collective on the
data path; one all-reduce of the timing scalar).

`--impl reference` times the reference's own CPU path (SciPy stft/istft + the
NumPy restatement of its packing / log / mask ops, oracle/ref_oracle.py) with all
host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 16000
WORKLOAD = dict(B=256, n=48000, N=512, H=128, S=3)
METRIC = "mixture audio-seconds per second, STFT->mask->iSTFT"
UNIT = "audio-s/s"


def frame_count(n, N, H):
    nadd = ((-n) % H) % N
    return (n + nadd) // H + 1, nadd


def algorithmic_bytes(B, n, N, H, S):
    """per mixture 4n + 4TN + 4n + 4*S*T*N/2 + 4*S*(T-1)*H."""
    T, _ = frame_count(n, N, H)
    stft_b = 4 * n + 4 * T * N
    synth_b = 4 * n + 4 * S * T * (N // 2) + 4 * S * (T - 1) * H
    return B * stft_b, B * synth_b


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.1] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------
# CPU reference arm
# --------------------------------------------------------------------------
def _cpu_make(seed, count, n, N, H, S):
    import numpy as np
    rng = np.random.default_rng(seed)
    T, _ = frame_count(n, N, H)
    x = (rng.standard_normal((count, n)) * 0.1).astype(np.float32)
    m = rng.random((count, S, T, N // 2), dtype=np.float32)
    return x, m


def _cpu_one(args):
    from oracle import ref_oracle as R
    x, m, N, H = args
    R.separate_utterance_scipy(x, m, N, H)
    return 1


def cpu_serial_baseline(budget_s=12.0, max_utts=4096):
    """1 core, serial per-utterance loop - how the reference drives SciPy
    (process.py:89, main.py:769-771)."""
    w = WORKLOAD
    x, m = _cpu_make(99, 16, w["n"], w["N"], w["H"], w["S"])
    _cpu_one((x[0], m[0], w["N"], w["H"]))            # warm SciPy's plan caches
    done, t0 = 0, time.perf_counter()
    while done < max_utts and time.perf_counter() - t0 < budget_s:
        _cpu_one((x[done % 16], m[done % 16], w["N"], w["H"]))
        done += 1
    dt = time.perf_counter() - t0
    return {"value": done * w["n"] / SR / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{done} utterance passes (16 distinct 3 s utterances of the {w['B']}-utterance workload, cycled), "
                      f"serial SciPy stft/istft x{w['S']} + NumPy pack/log/mask, {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU path on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    w = WORKLOAD
    cores = os.cpu_count() or 1
    probe = min(w["B"], max(cores, 8))
    x, m = _cpu_make(7, w["B"] if w["B"] <= 64 else 64, w["n"], w["N"], w["H"], w["S"])
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        mk = lambda cnt: [(x[i % len(x)], m[i % len(x)], w["N"], w["H"]) for i in range(cnt)]
        pool.map(_cpu_one, mk(probe), chunksize=1)                 # warm caches / imports in the workers
        tp = time.perf_counter()
        pool.map(_cpu_one, mk(probe), chunksize=1)
        per_utt = (time.perf_counter() - tp) / probe               # wall seconds per utterance with all cores busy
        # bounded sample: the whole --steps/--warmup run stays near 90 s
        per_step = int(90.0 / max(args.steps + args.warmup, 1) / max(per_utt, 1e-6))
        per_step = max(cores, min(w["B"], per_step))
        jobs = mk(per_step)
        for _ in range(args.warmup):
            pool.map(_cpu_one, jobs, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_one, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    val = args.steps * per_step * w["n"] / SR / dt
    sample = (f"{per_step} of {w['B']} utterances per step, {cores} worker processes; SciPy {__import__('scipy').__version__} "
              f"stft/istft as main.py:97/111 + NumPy restatement of utils.py/ops.py (TF 1.x not installable)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 TIMIT-shaped 256 x 3 s @16 kHz, FFT 512 hop 128, S=3 (bounded sample per step)", **w,
                   "sample_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from gan_sass_tf_b200 import _native
    from gan_sass_tf_b200.app import ops
    from gan_sass_tf_b200.app.spectral import SpectralPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # one process per GPU: keep the rank (and the pinned host buffers it allocates) on the CPUs / memory next to
        # its GPU, otherwise the ranks' host copies all cross the same socket link
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device() if os.environ.get("CUDA_VISIBLE_DEVICES") is None else local))
        except Exception:
            pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _native.lib()

    w = WORKLOAD
    B, n, N, H, S = w["B"], w["n"], w["N"], w["H"], w["S"]
    T, _ = frame_count(n, N, H)
    L = (T - 1) * H
    NSETS = 3                                           # rotating input sets: no L2 reuse between steps
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    waves = [(torch.randn(B, n, device=dev, generator=g) * 0.1).clamp_(-1, 1) for _ in range(NSETS)]
    masks = [torch.rand(B, S, T, N // 2, device=dev, generator=g) for _ in range(NSETS)]
    feat = torch.empty(B, T, N, device=dev)
    out = torch.empty(B * S, L, device=dev)
    stream = torch.cuda.current_stream()

    def step(i):
        k = i % NSETS
        lib = _native.lib()
        _native.check(lib.gss_stft_packed(waves[k].data_ptr(), B, n, n, N, H, _native.FLAG_LOG, 1e-7,
                                          feat.data_ptr(), stream.cuda_stream))
        _native.check(lib.gss_mask_istft(waves[k].data_ptr(), masks[k].data_ptr(), B, S, n, n, N, H,
                                         out.data_ptr(), L, stream.cuda_stream))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.15)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    lib = _native.lib()
    barrier()
    l0 = _native.launch_count()
    t0 = time.time()
    for i in range(args.steps):
        k = i % NSETS
        ev[i][0].record(stream)
        _native.check(lib.gss_stft_packed(waves[k].data_ptr(), B, n, n, N, H, _native.FLAG_LOG, 1e-7,
                                          feat.data_ptr(), stream.cuda_stream))
        ev[i][1].record(stream)
        _native.check(lib.gss_mask_istft(waves[k].data_ptr(), masks[k].data_ptr(), B, S, n, n, N, H,
                                         out.data_ptr(), L, stream.cuda_stream))
        ev[i][2].record(stream)
    barrier()
    t1 = time.time()
    launches = _native.launch_count() - l0
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    stft_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    synth_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    clocks = sampler.stop(t0, t1) if sampler else None

    # ---- e2e through the host-buffer API --------------------------------------
    # Every step uploads its mixture batch from pinned host memory and downloads its S separated
    # waveforms into pinned host memory; two workspace slots keep the upload of step i+1 and the
    # download of step i on the link at the same time (what a loop over many batches does).
    pipe = SpectralPipeline(B, n, S, N, H, device=dev, chunks=8, depth=2)
    for hbuf in pipe.wave_hs:
        hbuf.copy_(waves[0].cpu())
    e2e_steps = max(4, min(args.steps, 20))
    e2e_ms = float("nan")
    e2e16_ms = float("nan")
    pipe_f32_bytes = (pipe.h2d_bytes, pipe.d2h_bytes)
    e2e_checksum = 0.0
    if not args.no_e2e:
        def e2e_run(steps):
            acc = 0.0
            for i in range(steps):
                slot = i % pipe.depth
                if i >= pipe.depth:
                    acc += float(pipe.wait(slot)[0, 0])     # host-side read of the step's result (step i - depth)
                pipe.analyse(slot=slot, block=False)        # pinned host waves -> device -> log features (separator input)
                pipe.synthesise(masks[i % NSETS], slot=slot, block=False)   # masks (device) -> pinned host waveforms
            for hbuf in pipe.wait():
                acc += float(hbuf[0, 0])
            return acc
        e2e_run(max(2, min(args.warmup, 4)))
        barrier()
        tw0 = time.perf_counter()
        e2e_checksum = e2e_run(e2e_steps)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - tw0) * 1e3
        # the same loop with int16 PCM on the host link (the WAV sample format the reference reads and writes:
        # main.py:83, :112-116): half the bytes each way.  Reported beside e2e, not instead of it.
        del pipe
        pipe16 = SpectralPipeline(B, n, S, N, H, device=dev, chunks=8, depth=2, pcm16=True)
        for hbuf in pipe16.wave_hs:
            hbuf.copy_((waves[0] * 32767.0).to(torch.int16).cpu())
        pipe, pipe_f32_bytes = pipe16, (B * n * 4, B * S * L * 4)
        e2e_run(max(2, min(args.warmup, 4)))
        barrier()
        tw0 = time.perf_counter()
        e2e_run(e2e_steps)
        torch.cuda.synchronize()
        e2e16_ms = (time.perf_counter() - tw0) * 1e3

    # ---- untimed full-size property check + the metric all-reduce (the only collective) ----
    from gan_sass_tf_b200.app import parallel
    nchk = 16
    mk = masks[0][:nchk] + 0.05
    mk = (mk / mk.sum(dim=1, keepdim=True)).contiguous()
    rec = ops.mask_istft(waves[0][:nchk].contiguous(), mk, N, H).reshape(nchk, S, -1).sum(dim=1)[:, :n]
    snr = ops.batch_snr(waves[0][:nchk].contiguous(), rec.contiguous())          # dB per utterance (app/ops.py:162-189, EPS inside the logs caps it near 50 dB)
    xw = waves[0][:nchk].double()
    true_snr_db = float(10.0 * torch.log10(xw.pow(2).sum() / (xw - rec.double()).pow(2).sum().clamp_min(1e-300)))
    vec = parallel.metric_vector(float(snr.sum()), 0.0, float(snr.sum()), float(nchk), device=dev)
    recon_snr_db, _, _, checked = parallel.allreduce_metrics(vec)

    t = torch.tensor([total_ms, e2e_ms, stft_ms, synth_ms, e2e16_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, stft_ms, synth_ms, e2e16_ms = (float(v) for v in t.tolist())

    if rank == 0:
        audio_s = B * n / SR
        value = world * args.steps * audio_s / (total_ms * 1e-3)
        e2e_val = None if e2e_ms != e2e_ms else world * e2e_steps * audio_s / (e2e_ms * 1e-3)
        stft_b, synth_b = algorithmic_bytes(B, n, N, H, S)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = synth_b / (synth_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("mask_istft_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2 TIMIT-shaped 256 x 3 s @16 kHz, FFT 512 hop 128, S=3, per GPU", **w, "T": T,
                       "l2": "inputs larger than L2 (788 MB/step) and 3 rotating input sets",
                       "separator": "stand-in: U(0,1) masks resident in HBM", "parallelism": f"utterance-sharded x{world}"},
            "roofline": {"bound": "hbm", "kernel": "mask_istft_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "bytes_per_launch": synth_b, "ms_per_launch": synth_ms,
                         "limiter": "not HBM at this batch: issue cadence of the SM sub-partitions (half-rate FP32x2 + shared-memory "
                                    "instructions; static schedule 3779 cycles per frame pair and warp = 172 us per launch), "
                                    "DESIGN.md 4.5 / profiles/r1c_tmem_variant.txt",
                         "step_frac_of_hbm": (stft_b + synth_b) / (total_ms / args.steps * 1e-3) / 1e9 / peak,
                         "stft_kernel": {"bytes_per_launch": stft_b, "ms_per_launch": stft_ms,
                                         "achieved": stft_b / (stft_ms * 1e-3) / 1e9}},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": pipe_f32_bytes[0], "d2h_bytes_per_step": pipe_f32_bytes[1],
                    "steps": e2e_steps, "ms_per_step": None if e2e_ms != e2e_ms else e2e_ms / e2e_steps,
                    "api": "SpectralPipeline.analyse/synthesise(block=False)/wait, depth 2 -> gss_stft_h2d_async / gss_mask_istft_d2h_async / gss_wait_host "
                           "(pinned host buffers; host clock between device syncs)"},
            "e2e_pcm16": None if e2e16_ms != e2e16_ms else {
                "value": world * e2e_steps * audio_s / (e2e16_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e16_ms / e2e_steps,
                "h2d_bytes_per_step": B * n * 2, "d2h_bytes_per_step": B * S * L * 2,
                "what": "same loop, int16 PCM in / per-clip normalised int16 PCM out (SpectralPipeline(pcm16=True))"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "check": {"recon_snr_db": recon_snr_db, "recon_snr_db_no_eps": true_snr_db, "e2e_checksum": e2e_checksum, "utterances": int(checked),
                      "what": "sum_s iSTFT(mask_s * STFT(x)) vs x with sum_s mask_s = 1, mean over ranks (NCCL all-reduce)"},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_serial_baseline()
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: everything else that writes to file descriptor 1 (NCCL prints its version
    banner there at the first collective, libraries print warnings) is sent to stderr for the whole run."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
