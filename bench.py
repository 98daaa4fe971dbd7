#!/usr/bin/env python
"""bench.py - mixture audio-seconds per second through STFT -> mask -> iSTFT.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--path feature|wave]

A "step" is one pass of the hot path over one batch of synthetic mixtures
(BASELINE.json config C2 per GPU: 256 x 3 s @ 16 kHz, FFT 512, hop 128, S = 3):

    lin, log   = stft_dual(wave)                  kernel 1: A1+A2+A3, one transform, both feature forms
    [separator stand-in: per-source masks already resident in HBM]
    waves, ae  = mask_istft_feature(lin, mask)    kernel 2: A7+A8 (+ the A12 auto-encoder partial, fused)
    vec4       = metric_finalise(ae)              kernel 3: the batch's metric vector
    N > 1:       NCCL all-reduce(vec4)            on a side stream, overlapping the next step

(`--path wave` times round 1's step instead: stft_log(wave) + mask_istft(wave, mask), which recomputes the mixture
spectrum from the waveform; both are reported under `paths`.)

`value` is device-timed (CUDA events, max over ranks) with inputs resident in HBM; `e2e` is the same work through the
host-buffer API (pinned host waves in, pinned host waveforms out, copies inside the timed region) and carries the
copy-only `link_ceiling` of the same buffers.  N > 1: one process per GPU (torchrun), every rank runs its own C2 batch
(weak scaling, no collective on the data path).  The `c4` block times BASELINE.json config 4 (8192 x 4 s, sharded by
utterance over the ranks: strong scaling); `sustained` repeats the step back to back for >= 2 s with its own clocks.

`--impl reference` times the reference's own CPU path (SciPy stft/istft + the NumPy restatement of its packing / log /
mask ops, oracle/ref_oracle.py) with all host cores on a bounded sample of the same workload.
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
C4 = dict(B=8192, n=64000, N=512, H=128, S=3)
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


def config_dict(world):
    """The SAME dict from both arms (the driver compares them)."""
    w = WORKLOAD
    T, _ = frame_count(w["n"], w["N"], w["H"])
    return {"workload": "C2 TIMIT-shaped 256 x 3 s @16 kHz, FFT 512 hop 128, S=3, per GPU", **w, "T": T,
            "l2": "inputs larger than L2 (788 MB/step) and 3 rotating input sets",
            "separator": "stand-in: U(0,1) masks resident in HBM", "parallelism": f"utterance-sharded x{world}"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown CPU"


def git_head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True, timeout=5).stdout.strip() or None
    except Exception:
        return None


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

    def window(self, t0, t1):
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.1] or [r for (_, r) in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


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
    return {"value": done * w["n"] / SR / dt, "unit": UNIT, "cores": 1, "kind": "port", "cpu": cpu_model(),
            "host_cores": os.cpu_count(),
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
    sample = (f"{per_step} of {w['B']} utterances per step, {cores} worker processes on {cpu_model()}; SciPy {__import__('scipy').__version__} "
              f"stft/istft as main.py:97/111 + NumPy restatement of utils.py/ops.py (TF 1.x not installable)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "cpu": cpu_model(), "sample": sample,
                         "sample_per_step": per_step},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------
class Step:
    """One rank's step over its own batch: buffers, the two kernel paths, and the metric all-reduce."""

    def __init__(self, torch, dist, w, dev, rank, world, nsets, seed):
        from gan_sass_tf_b200 import _native
        self.t, self.dist, self.nv, self.world, self.dev = torch, dist, _native, world, dev
        self.B, self.n, self.N, self.H, self.S = w["B"], w["n"], w["N"], w["H"], w["S"]
        self.T, _ = frame_count(self.n, self.N, self.H)
        self.L = (self.T - 1) * self.H
        self.nsets = nsets
        g = torch.Generator(device=dev).manual_seed(seed)
        B, n, T, N, S = self.B, self.n, self.T, self.N, self.S
        self.waves = [(torch.randn(B, n, device=dev, generator=g) * 0.1).clamp_(-1, 1) for _ in range(nsets)]
        self.masks = [torch.rand(B, S, T, N // 2, device=dev, generator=g) for _ in range(nsets)]
        self.lin = [torch.empty(B, T, N, device=dev) for _ in range(nsets)]
        self.feat = torch.empty(B, T, N, device=dev)
        self.out = torch.empty(B * S, self.L, device=dev)
        self.ae_rows = torch.zeros(B, device=dev)
        self.vec = [torch.zeros(4, device=dev) for _ in range(2)]
        self.stream = torch.cuda.current_stream()
        self.comm = torch.cuda.Stream(device=dev) if world > 1 else None
        self.ev_ready = [torch.cuda.Event() for _ in range(2)]
        self.ev_reduced = [torch.cuda.Event() for _ in range(2)]
        self.reduced_once = [False, False]
        self.lib = _native.lib()
        self.fused_metric = self.N in (256, 512)          # the fused auto-encoder partial exists in the register-exchange kernels

    def feature(self, i, ev=None):
        """stft_dual -> mask_istft_feature (+ fused AE partial) -> metric vector (-> all-reduce on the side stream)"""
        k, v = i % self.nsets, i % 2
        lib, st, nv = self.lib, self.stream.cuda_stream, self.nv
        B, n, N, H, S, T, L = self.B, self.n, self.N, self.H, self.S, self.T, self.L
        if ev:
            ev[0].record(self.stream)
        nv.check(lib.gss_stft_packed_dual(self.waves[k].data_ptr(), B, n, n, N, H, 1e-7, self.lin[k].data_ptr(), self.feat.data_ptr(), st))
        if ev:
            ev[1].record(self.stream)
        nv.check(lib.gss_mask_istft_feature_ae(self.lin[k].data_ptr(), self.masks[k].data_ptr(), B, S, T, N, H, nv.FLAG_REVERSE,
                                               self.out.data_ptr(), L, self.ae_rows.data_ptr() if self.fused_metric else None, st))
        if ev:
            ev[2].record(self.stream)
        if not self.fused_metric:
            return
        if self.comm is not None and self.reduced_once[v]:
            self.stream.wait_event(self.ev_reduced[v])             # the all-reduce of step i-2 has consumed this vector
        nv.check(lib.gss_metric_finalise(self.ae_rows.data_ptr(), None, B, 1, 1, float(T * N), self.vec[v].data_ptr(), st))
        if self.comm is not None:
            self.ev_ready[v].record(self.stream)
            with self.t.cuda.stream(self.comm):
                self.comm.wait_event(self.ev_ready[v])
                work = self.dist.all_reduce(self.vec[v], op=self.dist.ReduceOp.SUM, async_op=True)
                work.wait()                                        # orders the side stream after NCCL's; the host does not block
                self.ev_reduced[v].record(self.comm)
            self.reduced_once[v] = True

    def wave(self, i, ev=None):
        """round 1's step: stft_log + mask_istft recomputing the mixture spectrum from the waveform (no fused metric)"""
        k = i % self.nsets
        lib, st, nv = self.lib, self.stream.cuda_stream, self.nv
        B, n, N, H, S, L = self.B, self.n, self.N, self.H, self.S, self.L
        if ev:
            ev[0].record(self.stream)
        nv.check(lib.gss_stft_packed(self.waves[k].data_ptr(), B, n, n, N, H, nv.FLAG_LOG, 1e-7, self.feat.data_ptr(), st))
        if ev:
            ev[1].record(self.stream)
        nv.check(lib.gss_mask_istft(self.waves[k].data_ptr(), self.masks[k].data_ptr(), B, S, n, n, N, H, self.out.data_ptr(), L, st))
        if ev:
            ev[2].record(self.stream)

    def join(self):
        """the compute stream waits for every outstanding all-reduce (inside the timed region)"""
        if self.comm is not None:
            for v in range(2):
                if self.reduced_once[v]:
                    self.stream.wait_event(self.ev_reduced[v])


def run_native(args):
    import torch
    import torch.distributed as dist
    from gan_sass_tf_b200 import _native
    from gan_sass_tf_b200.app import ops, parallel
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
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        except Exception:
            pass
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    _native.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = WORKLOAD
    B, n, N, H, S = w["B"], w["n"], w["N"], w["H"], w["S"]
    T, _ = frame_count(n, N, H)
    L = (T - 1) * H
    NSETS = 3                                           # rotating input sets: no L2 reuse between steps
    sp = Step(torch, dist, w, dev, rank, world, NSETS, 1234 + rank)
    stream = sp.stream
    primary = sp.feature if args.path == "feature" else sp.wave
    other = sp.wave if args.path == "feature" else sp.feature

    def timed(fn, steps, warmup, per_kernel=True):
        for i in range(warmup):
            fn(i)
        sp.join()
        barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)] if per_kernel else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = _native.launch_count()
        t0 = time.time()
        e0.record(stream)
        for i in range(steps):
            fn(i, ev[i] if ev else None)
        sp.join()
        e1.record(stream)
        barrier()
        t1 = time.time()
        total = e0.elapsed_time(e1)
        k1 = sum(e[0].elapsed_time(e[1]) for e in ev) / steps if ev else None
        k2 = sum(e[1].elapsed_time(e[2]) for e in ev) / steps if ev else None
        return total, k1, k2, _native.launch_count() - l0, (t0, t1)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.15)
    total_ms, stft_ms, synth_ms, launches, (t0, t1) = timed(primary, args.steps, args.warmup)
    clocks = sampler.window(t0, t1) if sampler else None
    ae_vec = sp.vec[(args.steps - 1) % 2].clone() if args.path == "feature" else None
    alt_ms, alt_k1, alt_k2, _, _ = timed(other, max(5, min(args.steps, 20)), 3)
    alt_steps = max(5, min(args.steps, 20))

    # ---- sustained: the same step back to back for >= 2 s, with its own clocks record ------------------
    sustained = None
    if not args.no_sustained:
        tper = torch.tensor([total_ms / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tper, op=dist.ReduceOp.MAX)      # every rank must run the same number of steps (one collective each)
        ns = int(2200.0 / max(float(tper[0]), 1e-3)) + 1
        s_ms, _, _, _, (s0, s1) = timed(primary, ns, 3, per_kernel=False)
        sustained = {"steps": ns, "seconds": s_ms * 1e-3, "ms_per_step": s_ms / ns, "clocks": sampler.window(s0, s1) if sampler else None}

    # ---- e2e through the host-buffer API --------------------------------------
    # Every step uploads its mixture batch from pinned host memory and downloads its S separated
    # waveforms into pinned host memory; two workspace slots keep the upload of step i+1 and the
    # download of step i on the link at the same time (what a loop over many batches does).
    e2e_steps = max(4, min(args.steps, 20))
    e2e_ms = e2e16_ms = link_ms = float("nan")
    e2e_checksum = 0.0
    if not args.no_e2e:
        pipe = SpectralPipeline(B, n, S, N, H, device=dev, chunks=args.chunks, depth=2)
        for hbuf in pipe.wave_hs:
            hbuf.copy_(sp.waves[0].cpu())

        def e2e_run(pipe, steps):
            acc = 0.0
            for i in range(steps):
                slot = i % pipe.depth
                if i >= pipe.depth:
                    acc += float(pipe.wait(slot)[0, 0])     # host-side read of the step's result (step i - depth)
                pipe.analyse(slot=slot, block=False)        # pinned host waves -> device -> log features (separator input)
                pipe.synthesise(sp.masks[i % NSETS], slot=slot, block=False)   # masks (device) -> pinned host waveforms
            for hbuf in pipe.wait():
                acc += float(hbuf[0, 0])
            return acc

        def e2e_timed(pipe):
            e2e_run(pipe, max(2, min(args.warmup, 4)))
            barrier()
            tw0 = time.perf_counter()
            chk = e2e_run(pipe, e2e_steps)
            torch.cuda.synchronize()
            return (time.perf_counter() - tw0) * 1e3, chk

        e2e_ms, e2e_checksum = e2e_timed(pipe)
        # copy-only ceiling: the same pinned buffers and bytes, both directions at once, no kernels, every rank
        # at the same time (what the host link gives this process while its peers do the same)
        s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def link_run(steps):
            for i in range(steps):
                slot = i % pipe.depth
                with torch.cuda.stream(s_up):
                    pipe.wave_ds[slot].copy_(pipe.wave_hs[slot], non_blocking=True)
                with torch.cuda.stream(s_dn):
                    pipe.out_hs[slot].copy_(pipe.out_ds[slot], non_blocking=True)
            s_up.synchronize(); s_dn.synchronize()
        link_run(3)
        barrier()
        tw0 = time.perf_counter()
        link_run(e2e_steps)
        link_ms = (time.perf_counter() - tw0) * 1e3
        del pipe
        # the same loop with int16 PCM on the host link (the WAV sample format the reference reads and writes:
        # main.py:83, :112-116): half the bytes each way.  Reported beside e2e, not instead of it.
        pipe16 = SpectralPipeline(B, n, S, N, H, device=dev, chunks=args.chunks, depth=2, pcm16=True)
        for hbuf in pipe16.wave_hs:
            hbuf.copy_((sp.waves[0] * 32767.0).to(torch.int16).cpu())
        e2e16_ms, _ = e2e_timed(pipe16)
        del pipe16

    # ---- untimed full-size property check (round trip through the timed path's kernels) ----
    nchk = 16
    mk = sp.masks[0][:nchk] + 0.05
    mk = (mk / mk.sum(dim=1, keepdim=True)).contiguous()
    xs = sp.waves[0][:nchk].contiguous()
    if args.path == "feature":
        rec = ops.mask_istft_feature(ops.stft(xs, N, H), mk, H).reshape(nchk, S, -1).sum(dim=1)[:, :n]
    else:
        rec = ops.mask_istft(xs, mk, N, H).reshape(nchk, S, -1).sum(dim=1)[:, :n]
    xw = xs.double()
    true_snr_db = float(10.0 * torch.log10(xw.pow(2).sum() / (xw - rec.double()).pow(2).sum().clamp_min(1e-300)))
    ae_mean = None
    if ae_vec is not None:
        ae_mean = float(ae_vec[1] / ae_vec[3])           # after the all-reduce: mean over every rank's mixtures
        ae_count = float(ae_vec[3])

    # ---- C4: 8192 x 4 s sharded by utterance over the ranks (strong scaling) ----------------------------
    c4 = None
    if not args.no_c4:
        del sp.waves, sp.masks, sp.lin, sp.feat, sp.out
        torch.cuda.empty_cache()
        lo, hi = parallel.shard_range(C4["B"], rank, world)
        wc = dict(C4, B=hi - lo)
        sc = Step(torch, dist, wc, dev, rank, world, 1, 4321 + rank)
        sp, keep = sc, sp                                 # `timed` drives sp
        fn = sc.feature if args.path == "feature" else sc.wave
        c4_steps = 5
        c4_ms, c4_k1, c4_k2, _, _ = timed(fn, c4_steps, 2)
        sp = keep
        tt = torch.tensor([c4_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        c4_ms = float(tt[0])
        sb, yb = algorithmic_bytes(C4["B"], C4["n"], C4["N"], C4["H"], C4["S"])
        c4 = {"workload": "C4 8192 x 4 s @16 kHz, FFT 512 hop 128, S=3, utterance-sharded over the ranks (shard_range)",
              **C4, "rows_on_rank0": hi - lo, "steps": c4_steps, "ms_per_step": c4_ms / c4_steps,
              "value": C4["B"] * C4["n"] / SR / (c4_ms / c4_steps * 1e-3), "unit": UNIT, "scaling": "strong",
              "algorithmic_bytes": sb + yb, "resident_gb_rank0": torch.cuda.max_memory_allocated(dev) / 1e9}
        del sc
        torch.cuda.empty_cache()

    # ---- the collective alone: latency of one 16-byte all-reduce, back to back on the compute stream (N > 1) ----
    coll_us = None
    if world > 1:
        v4 = torch.zeros(4, device=dev)
        for _ in range(5):
            dist.all_reduce(v4)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = 200
        c0.record(stream)
        for _ in range(nrep):
            dist.all_reduce(v4)
        c1.record(stream)
        barrier()
        tc = torch.tensor([c0.elapsed_time(c1) / nrep * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        coll_us = float(tc[0])

    # ---- C5 (FFT_SIZE sweep, B = 1024 x 3 s, hop N/4) and C3 (one 60 s clip, FFT 1024 / hop 256): one GPU only ----
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = {"what": "BASELINE configs 5 and 3 through the same step (stft_dual + mask_istft_feature, S = 3; the fused metric "
                         "exists for 256 / 512 only), device-resident, 10 steps after 3 warm-ups each", "unit": UNIT, "c5": {}, "c3": None}
        peak_gbs = 6552.6
        try:
            peak_gbs = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak_gbs))
        except Exception:
            pass
        cases = [("c5", dict(B=1024, n=48000, N=Nf, H=Nf // 4, S=3)) for Nf in (256, 512, 1024, 2048, 4096)]
        cases.append(("c3", dict(B=1, n=960000, N=1024, H=256, S=3)))
        keep = sp
        for tag, wc in cases:
            sp = Step(torch, dist, wc, dev, rank, world, 2, 77)
            ms, k1, k2, _, _ = timed(sp.feature, 10, 3)
            sb, yb = algorithmic_bytes(wc["B"], wc["n"], wc["N"], wc["H"], wc["S"])
            row = {"ms_per_step": ms / 10, "analysis_ms": k1, "synthesis_ms": k2, "value": wc["B"] * wc["n"] / SR / (ms / 10 * 1e-3),
                   "step_frac_of_hbm": (sb + yb) / (ms / 10 * 1e-3) / 1e9 / peak_gbs, "synthesis_frac_of_hbm": yb / (k2 * 1e-3) / 1e9 / peak_gbs}
            if tag == "c5":
                sweep["c5"][str(wc["N"])] = row
            else:
                sweep["c3"] = dict(row, us_per_clip=ms / 10 * 1e3, note="one clip: launch- and latency-bound (58 MB of algorithmic bytes)")
            del sp
            torch.cuda.empty_cache()
        sp = keep

    t = torch.tensor([total_ms, e2e_ms, stft_ms, synth_ms, e2e16_ms, alt_ms, link_ms,
                      sustained["seconds"] * 1e3 if sustained else float("nan")], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, stft_ms, synth_ms, e2e16_ms, alt_ms, link_ms, sus_ms = (float(v) for v in t.tolist())

    if sampler:
        sampler.stop()
    if rank == 0:
        audio_s = B * n / SR
        value = world * args.steps * audio_s / (total_ms * 1e-3)
        nan = lambda v: v != v
        e2e_val = None if nan(e2e_ms) else world * e2e_steps * audio_s / (e2e_ms * 1e-3)
        stft_b, synth_b = algorithmic_bytes(B, n, N, H, S)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = synth_b / (synth_ms * 1e-3) / 1e9
        traffic, traffic_note = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            key = "mask_istft_feature_bytes_per_launch" if args.path == "feature" else "mask_istft_bytes_per_launch"
            traffic = tj.get(key)
            traffic_note = {"source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
                            "measured_at_git_head": tj.get("git_head"), "bench_git_head": git_head(),
                            "note": tj.get("note")}
        except Exception:
            pass
        kname = "mask_istft_kernel<512,2,3,4,FEAT,AE> (gss_mask_istft_feature_ae)" if args.path == "feature" else "mask_istft_kernel<512,2,3,4> (gss_mask_istft)"
        alt_name = "wave" if args.path == "feature" else "feature"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(world),
            "path": args.path,
            "paths": {args.path: {"ms_per_step": total_ms / args.steps, "analysis_ms": stft_ms, "synthesis_ms": synth_ms},
                      alt_name: {"ms_per_step": alt_ms / alt_steps, "analysis_ms": alt_k1, "synthesis_ms": alt_k2, "steps": alt_steps},
                      "what": "feature = stft_dual + mask_istft_feature (+ fused AE partial, metric vector, all-reduce at N > 1); "
                              "wave = stft_log + mask_istft recomputing the mixture spectrum (round 1's step, no metric)"},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_info": traffic_note,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "bytes_per_launch": synth_b, "ms_per_launch": synth_ms,
                         "bytes_note": "algorithmic bytes of SURVEY 8(d) (wave re-read + masks + outputs); the feature-fed kernel really reads "
                                       "the 4TN-byte spectrum instead of the 4n-byte waveform, see traffic",
                         "limiter": "issue cadence of the SM sub-partitions (half-rate FP32x2 + shared-memory instructions), DESIGN.md 4.5",
                         "step_frac_of_hbm": (stft_b + synth_b) / (total_ms / args.steps * 1e-3) / 1e9 / peak,
                         "stft_kernel": {"bytes_per_launch": stft_b, "ms_per_launch": stft_ms,
                                         "achieved": stft_b / (stft_ms * 1e-3) / 1e9}},
            "sustained": None if sustained is None else {
                "value": world * sustained["steps"] * audio_s / (sus_ms * 1e-3), "unit": UNIT, "steps": sustained["steps"],
                "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / sustained["steps"], "clocks": sustained["clocks"]},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * n * 4, "d2h_bytes_per_step": B * S * L * 4,
                    "steps": e2e_steps, "ms_per_step": None if nan(e2e_ms) else e2e_ms / e2e_steps,
                    "link_ceiling": None if nan(link_ms) else {
                        "value": world * e2e_steps * audio_s / (link_ms * 1e-3), "unit": UNIT, "ms_per_step": link_ms / e2e_steps,
                        "frac_of_ceiling": None if nan(e2e_ms) else link_ms / e2e_ms,
                        "what": "the same pinned buffers and bytes copied both ways at once with no kernels, all ranks concurrently"},
                    "chunks": args.chunks,
                    "api": "SpectralPipeline.analyse/synthesise(block=False)/wait, depth 2 -> gss_stft_h2d_async / gss_mask_istft_d2h_async / gss_wait_host "
                           "(pinned host buffers; host clock between device syncs)"},
            "e2e_pcm16": None if nan(e2e16_ms) else {
                "value": world * e2e_steps * audio_s / (e2e16_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e16_ms / e2e_steps,
                "h2d_bytes_per_step": B * n * 2, "d2h_bytes_per_step": B * S * L * 2,
                "what": "same loop, int16 PCM in / per-clip normalised int16 PCM out (SpectralPipeline(pcm16=True))"},
            "c4": c4,
            "sweep": sweep,
            "gpu_launches": int(launches),
            "collective": None if world == 1 else {
                "what": "one NCCL all-reduce (sum) of the 4-float metric vector per step, on a side stream behind the step's metric kernel, "
                        "overlapping the next step; the timed region ends after the last one has landed", "per_step": 1,
                "allreduce_latency_us": coll_us,
                "latency_what": "one 16-byte NCCL all-reduce, 200 back to back on the compute stream, CUDA events, max over ranks: what "
                                "every step would pay if the collective were serialised with the kernels instead of hidden under the next step"},
            "clocks": clocks,
            "check": {"recon_snr_db": true_snr_db, "ae_loss_mean": ae_mean, "ae_count": ae_count if ae_vec is not None else None,
                      "e2e_checksum": e2e_checksum, "utterances": nchk,
                      "what": "sum_s iSTFT(mask_s * STFT(x)) vs x with sum_s mask_s = 1 (dB); ae_loss_mean = the timed steps' fused "
                              "auto-encoder partial after the all-reduce (mean over all ranks' mixtures, main.py:353-361)"},
        }
        if c4 is not None:
            c4["roofline_frac"] = c4["algorithmic_bytes"] / (c4["ms_per_step"] * 1e-3) / 1e9 / peak / world
            c4["per_gpu_value"] = c4["value"] / world
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
    ap.add_argument("--path", default="feature", choices=["feature", "wave"],
                    help="feature: stft_dual + mask_istft_feature (default); wave: stft_log + mask_istft (round 1's step)")
    ap.add_argument("--chunks", type=int, default=4, help="host-pipeline chunks per batch (e2e leg)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (profiling runs)")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 (8192 x 4 s) block")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the C5 / C3 block (one-GPU runs only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
