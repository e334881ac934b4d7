#!/usr/bin/env python
"""Benchmark of the hot path: layer-0 HM synthesis frames/s on BASELINE.json configs[1]
(batch = 1024 synthetic 2-s utterances, 44.1 kHz, 5 ms hop, 128 harmonics) per GPU.

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                   (the reference's own CPU path, host cores)

One "step" = one llsm_synthesize pass (y_sin, y_noise, y) over the whole batch. Prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "layer0 analysis+synthesis frames/sec @44.1kHz/5ms hop; HBM GB/s vs roofline"
UNIT = "frames/s"
BANK_BYTES_PER_FRAME = 1910.0     # 1028 B parameters in + 882 B of y_sin out
FULL_BYTES_PER_FRAME = 5886.0     # whole HM synthesis frame


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="utterances per GPU")
    ap.add_argument("--nfrm", type=int, default=400)
    ap.add_argument("--nhar", type=int, default=128)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-analysis", action="store_true")
    ap.add_argument("--analysis-batch", type=int, default=256, help="utterances of the analysis leg")
    return ap.parse_args()


def workload(args, distinct=32, seed=0):
    from libllsm2_b200.synthetic import synth_frames
    d = min(distinct, args.batch)
    fr, conf = synth_frames(d, args.nfrm, nhar=args.nhar, seed=seed)
    reps = (args.batch + d - 1) // d
    full = {}
    for k, v in fr.items():
        full[k] = None if v is None else np.ascontiguousarray(np.concatenate([v] * reps, 0)[:args.batch])
    conf.nutt = args.batch
    return full, conf, fr


# ------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------ reference (CPU) arm
def _ref_lib(fast=True):
    name = "libllsm2_ref_fast.so" if fast else "libllsm2_ref.so"
    path = os.path.join(ROOT, "oracle", "_ref", name)
    lib = C.CDLL(path)
    lib.ref_time_synthesize_soa.restype = C.c_double
    return lib


_G = {}   # fork-inherited workload of the reference arm (nothing is pickled per step)


def _ref_time_utts(job):
    """Worker: time llsm_synthesize (reference sources, -Ofast) on a list of utterances; returns
    (seconds inside llsm_synthesize, frames)."""
    fr, conf_t, idxs = job
    if fr is None:
        fr = _G["frames"]
    lib = _ref_lib(True)
    (nfrm, fs, thop, maxnhar, maxnhar_e, npsd, nch, cf, lip) = conf_t
    cfa = np.array(cf, np.float32)
    tot = 0.0
    for b in idxs:
        a = [np.ascontiguousarray(fr[k][b]) for k in
             ("f0", "nhar", "ampl", "phse", "psd", "psdres", "edc", "enhar", "eampl", "ephse")]
        tot += lib.ref_time_synthesize_soa(1, nfrm, C.c_float(fs), C.c_float(thop), maxnhar, maxnhar_e,
                                           npsd, nch, cfa.ctypes.data_as(C.c_void_p), C.c_float(lip), 1,
                                           *[x.ctypes.data_as(C.c_void_p) for x in a])
    return tot, nfrm * len(idxs)


def _conf_tuple(conf):
    return (conf.nfrm, conf.fs, conf.thop, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel,
            list(conf.chanfreq), conf.lip_radius)


def cpu_baseline_single(distinct, conf, nutt=8):
    """Reference C path (oracle/_ref, the reference's Release flags), ONE thread, bounded sample."""
    idxs = list(range(min(nutt, len(distinct["f0"]))))
    _ref_time_utts((distinct, _conf_tuple(conf), idxs[:1]))     # warm-up
    t, frames = _ref_time_utts((distinct, _conf_tuple(conf), idxs))
    return {"value": frames / t, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": "%d utterances x %d frames of the bench workload, llsm_synthesize of the "
                      "unmodified reference sources + ciglet shim, gcc -Ofast, single thread" % (len(idxs), conf.nfrm)}


def cpu_baseline_analysis(x, distinct, conf, nutt=2):
    """llsm_analyze of the reference sources (gcc -Ofast), one thread, on nutt synthesised utterances."""
    lib = _ref_lib(True)
    lib.ref_time_analyze.restype = C.c_double
    cf = np.array(list(conf.chanfreq), np.float32)
    tot = 0.0
    for b in range(min(nutt, x.shape[0])):
        xb = np.ascontiguousarray(x[b], np.float32); f0 = np.ascontiguousarray(distinct["f0"][b], np.float32)
        tot += lib.ref_time_analyze(1, xb.ctypes.data_as(C.c_void_p), int(xb.shape[0]), C.c_float(conf.fs),
                                    f0.ctypes.data_as(C.c_void_p), conf.nfrm, C.c_float(conf.thop), conf.maxnhar,
                                    conf.maxnhar_e, conf.npsd, conf.nchannel, cf.ctypes.data_as(C.c_void_p), 1)
    n = min(nutt, x.shape[0])
    return {"value": n * conf.nfrm / tot, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": "%d utterances x %d frames, llsm_analyze (CZT) of the reference sources, gcc -Ofast, single thread" % (n, conf.nfrm)}


def usable_cores():
    """Host threads this process may really use: affinity mask capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, q // per))
        except (OSError, ValueError, IndexError):
            pass
    return max(1, n)


def run_reference(args):
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    full, conf, distinct = workload(args, distinct=16)
    ncore = usable_cores()
    per_core = 8
    _G["frames"] = distinct
    pool = mp.get_context("fork").Pool(ncore)
    jobs = [(None, _conf_tuple(conf), [(c * per_core + i) % 16 for i in range(per_core)]) for c in range(ncore)]
    step_frames = conf.nfrm * per_core * ncore
    for _ in range(args.warmup):
        pool.map(_ref_time_utts, jobs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.map(_ref_time_utts, jobs)
    dt = time.perf_counter() - t0
    pool.close()
    v = step_frames * args.steps / dt
    sample = ("each step = %d utterances x %d frames (8 per host core) of the bench workload through "
              "llsm_synthesize of the unmodified reference sources + ciglet shim (gcc -Ofast), one process "
              "per core" % (per_core * ncore, conf.nfrm))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(args, conf),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncore, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def config_dict(args, conf):
    return {"workload": "BASELINE configs[1]: batch=%d synthetic 2 s utterances per GPU, 44.1 kHz, 5 ms hop "
                        "(%d frames), %d harmonics, layer-0 HM synthesis (y_sin + y_noise + y)"
                        % (args.batch, conf.nfrm, args.nhar),
            "batch_per_gpu": args.batch, "nfrm": conf.nfrm, "nhar": args.nhar, "npsd": conf.npsd,
            "nchannel": conf.nchannel, "fs": conf.fs, "thop": conf.thop,
            "noise": "device Philox templates drawn inside the timed step",
            "l2": "inputs+outputs per step (~3 GB) exceed the 126 MB L2; no explicit flush",
            "parallelism": "utterance shards, no data-path collective"}


# ------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import libllsm2_b200 as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    full, conf, distinct = workload(args, seed=rank)
    ctx = L.Context(local)
    d = {k: (torch.from_numpy(v).to(dev) if v is not None else None) for k, v in full.items()}
    ny = L.output_length(conf.nfrm, conf.thop, conf.fs)
    out = {k: torch.empty((conf.nutt, ny), dtype=torch.float32, device=dev) for k in ("y", "y_sin", "y_noise")}
    frames_per_step = conf.nutt * conf.nfrm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        L.synthesize_l0(ctx, conf, d, white=None, seed=1000 + i, out=out)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None

    # the harmonic-bank kernel alone (the kernel BASELINE.json's roofline target names; tcgen05 path,
    # kernels_bank_tc.cuh), CUDA events on the launching stream
    ys = out["y_sin"]
    for i in range(args.warmup):
        L.synthesize_harmonics(ctx, conf, d, ny, out=ys)
    torch.cuda.synchronize()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for i in range(args.steps):
        L.synthesize_harmonics(ctx, conf, d, ny, out=ys)
    b1.record()
    torch.cuda.synchronize()
    bank_ms = b0.elapsed_time(b1) / args.steps

    # every kernel of the step, live: CUDA events recorded by the library between its launches (a separate pass, so
    # that the timed region above stays free of them)
    ctx.set_kernel_timing(True)
    kms = {}
    for i in range(max(3, min(args.steps, 10))):
        step(10_000 + i)
        for k, v in ctx.kernel_times().items():
            kms.setdefault(k, []).append(v)
    ctx.set_kernel_timing(False)
    kms = {k: float(np.median(v)) for k, v in kms.items()}

    # end to end through the host-buffer C ABI (pinned host memory, H2D + kernels + D2H per step)
    e2e = None
    if not args.no_e2e:
        pin = {k: (torch.from_numpy(v).pin_memory() if v is not None else None) for k, v in full.items()}
        hout = {k: torch.empty((conf.nutt, ny), dtype=torch.float32).pin_memory() for k in ("y", "y_sin", "y_noise")}
        h2d = sum(v.numel() * v.element_size() for v in pin.values() if v is not None)
        d2h = sum(v.numel() * v.element_size() for v in hout.values())
        for i in range(2):
            L.synthesize_l0_host(ctx, conf, pin, white=None, seed=i, out=hout)
        barrier()
        t0 = time.perf_counter()
        ne = max(2, min(args.steps, 5))
        for i in range(ne):
            L.synthesize_l0_host(ctx, conf, pin, white=None, seed=i, out=hout)
        barrier()
        dt = time.perf_counter() - t0
        e2e_t = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e = {"value": frames_per_step * world * ne / float(e2e_t.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": ne,
               "note": "llsm_b200_synthesize_l0_host: pinned host frames in, y/y_sin/y_noise out"}

    # analysis leg (llsm_analyze, CZT harmonics) on the waveforms the synthesis leg just produced
    ana = None
    if not args.no_analysis:
        nb = min(args.analysis_batch, conf.nutt)
        from libllsm2_b200 import abi
        ca = abi.make_conf(nb, conf.nfrm, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop,
                           list(conf.chanfreq)[:conf.nchannel - 1], conf.lip_radius)
        xa = out["y"][:nb].contiguous(); fa = d["f0"][:nb].contiguous()
        for _ in range(2):
            L.analyze_l0(ctx, ca, xa, fa)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        la = ctx.launches
        na = max(2, min(args.steps, 5))
        a0.record()
        for _ in range(na):
            L.analyze_l0(ctx, ca, xa, fa)
        a1.record(); torch.cuda.synchronize()
        ana_ms = a0.elapsed_time(a1) / na
        ana = {"value": nb * conf.nfrm * world / (ana_ms * 1e-3), "unit": UNIT, "batch_per_gpu": nb, "ms_per_call": ana_ms,
               "gpu_launches": int((ctx.launches - la) // na),
               "note": "llsm_b200_analyze_l0 (f0 refine, CZT harmonics, residual, noise PSD + Kalman/RTS, sub-band "
                       "envelopes) on %d of the synthesised utterances, device-resident" % nb}

    t = torch.tensor([ms, bank_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, bank_ms = float(t[0].item()), float(t[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = BANK_BYTES_PER_FRAME * frames_per_step / (bank_ms * 1e-3) / 1e9
    # algorithmic bytes per frame of the other kernels: shaper 2 048 B model PSDs + 882 B excitation in
    # + 882 B out; excitation 168 B envelope parameters + 882 B out; the template kernels only touch scratch
    # (B x nchannel templates of ~20 128 samples: written once by the fill, read + written by the IIR)
    tmpl_bytes = float(conf.nutt) * conf.nchannel * 20128 * 4
    kbytes = {"hm_bank": BANK_BYTES_PER_FRAME * frames_per_step, "noise_shape": 3812.0 * frames_per_step,
              "noise_excitation": 1050.0 * frames_per_step, "iir_filtfilt": 2 * tmpl_bytes, "white_fill": tmpl_bytes}
    kernels = {k: {"ms": v, "algorithmic_bytes_per_launch": kbytes[k], "achieved_gbs": kbytes[k] / (v * 1e-3) / 1e9,
                   "frac": kbytes[k] / (v * 1e-3) / 1e9 / peak} for k, v in kms.items()}
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "bank_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    value = frames_per_step * world * args.steps / (ms * 1e-3)
    res = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, conf), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "analysis": ana,
        "roofline": {"bound": "hbm", "kernel": "hm_bank_tc_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "ms_per_launch": bank_ms,
                     "algorithmic_bytes_per_launch": BANK_BYTES_PER_FRAME * frames_per_step,
                     "whole_step_gbs": FULL_BYTES_PER_FRAME * frames_per_step / (ms / args.steps * 1e-3) / 1e9},
        # the same accounting for every kernel of the step (CUDA events inside the library, llsm_b200_kernel_times);
        # "dominant" is the longest one -- the harmonic bank above is the kernel BASELINE.json's target names
        "kernels": kernels, "dominant_kernel": dom,
    }
    if ana is not None:
        syn = value
        ana["analysis_plus_synthesis"] = 1.0 / (1.0 / ana["value"] + 1.0 / syn)
    if not args.no_cpu_baseline and world == 1:
        try:
            res["cpu_baseline"] = cpu_baseline_single(distinct, conf)
            if ana is not None:
                ana["cpu_baseline"] = cpu_baseline_analysis(out["y"][:2].cpu().numpy(), distinct, conf)
        except Exception as e:  # the oracle .so travels with the repo; report rather than die
            res["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference",
                                   "sample": "unavailable: %s" % e}
    print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
