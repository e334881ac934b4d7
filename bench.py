#!/usr/bin/env python
"""Benchmark of the hot path: layer-0 analysis + synthesis frames/s (BASELINE.json's metric) on BASELINE
configs[1] shapes -- batch = 1024 synthetic 2-s utterances per GPU, 44.1 kHz, 5 ms hop, 128 harmonics.

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                   (the reference's own CPU path, all host cores)
  python bench.py --shard frames --gpus N ...            (frame-range sharding + the NCCL halo all-gather, synthesis)

One "step" = the chain of test/test-layer0-anasynth.c:40-46 over the whole batch: llsm_analyze of every waveform
(f0 refinement, CZT harmonics, residual, noise PSD + Kalman/RTS, sub-band envelopes) and llsm_synthesize of the chunk
it produced (harmonic bank, noise excitation, noise shaping, mix). `value` times it on device-resident buffers,
`e2e` through llsm_b200_anasynth_host with pinned host waveforms in and out. Prints ONE JSON line.
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
# SURVEY.md 8(d), algorithmic bytes per frame (f32, every array read / written once)
ANA_BYTES_PER_FRAME = 5130.0      # 882 B samples + f0 in; ampl, phse, noise model, f0 out
SYN_BYTES_PER_FRAME = 5886.0      # 3240 B parameters in; y, y_sin, y_noise out
BANK_BYTES_PER_FRAME = 1910.0     # 1028 B parameters in + 882 B of y_sin out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="utterances per GPU")
    ap.add_argument("--nfrm", type=int, default=400)
    ap.add_argument("--nhar", type=int, default=128)
    ap.add_argument("--distinct", type=int, default=256, help="distinct synthetic utterances (tiled to --batch)")
    ap.add_argument("--shard", default="utterances", choices=["utterances", "frames"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary legs (per-leg times, C1 arctic)")
    return ap.parse_args()


def workload(args, distinct=None, seed=0):
    from libllsm2_b200.synthetic import synth_frames
    d = min(distinct or args.distinct, args.batch)
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
_REF = {}


def _ref_lib():
    """oracle/_ref/libllsm2_ref_fast.so: the unmodified reference sources + ciglet shim, the reference's Release flags
    (-Ofast). Loaded once per process; the reference arm loads it in the PARENT so that the forked workers (and the
    driver's record of loaded libraries) share that mapping."""
    if "lib" not in _REF:
        lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libllsm2_ref_fast.so"))
        lib.ref_time_synthesize_soa.restype = C.c_double
        lib.ref_time_analyze.restype = C.c_double
        _REF["lib"] = lib
    return _REF["lib"]


_G = {}   # fork-inherited workload of the reference arm (nothing is pickled per step)


def _conf_tuple(conf):
    return (conf.nfrm, conf.fs, conf.thop, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel,
            list(conf.chanfreq), conf.lip_radius)


def _ref_synth_wave(fr, conf_t, b):
    """Reference llsm_synthesize of synthetic utterance b (untimed: it makes the analysis input)."""
    lib = _ref_lib()
    (nfrm, fs, thop, maxnhar, maxnhar_e, npsd, nch, cf, lip) = conf_t
    cfa = np.array(cf, np.float32)
    ny = lib.ref_output_length(nfrm, C.c_float(thop), C.c_float(fs))
    y = np.zeros(ny, np.float32); ys = np.zeros(ny, np.float32); yn = np.zeros(ny, np.float32)
    a = [np.ascontiguousarray(fr[k][b]) for k in
         ("f0", "nhar", "ampl", "phse", "psd", "psdres", "edc", "enhar", "eampl", "ephse")]
    lib.ref_synthesize_soa(nfrm, C.c_float(fs), C.c_float(thop), maxnhar, maxnhar_e, npsd, nch,
                           cfa.ctypes.data_as(C.c_void_p), C.c_float(lip), 1,
                           *[v.ctypes.data_as(C.c_void_p) for v in a], C.c_uint(1 + b),
                           y.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p), yn.ctypes.data_as(C.c_void_p))
    return y


def _ref_anasynth(job):
    """Worker: the reference chain llsm_analyze -> llsm_synthesize on a list of utterances.
    Returns (seconds in llsm_analyze, seconds in llsm_synthesize, frames)."""
    conf_t, idxs, hm = job
    lib = _ref_lib()
    (nfrm, fs, thop, maxnhar, maxnhar_e, npsd, nch, cf, lip) = conf_t
    cfa = np.array(cf, np.float32)
    t = (C.c_double * 2)(0.0, 0.0)
    for b in idxs:
        x = _G["x"][b]; f0 = _G["f0"][b]
        lib.ref_time_anasynth(1, x.ctypes.data_as(C.c_void_p), int(len(x)), C.c_float(fs), f0.ctypes.data_as(C.c_void_p),
                              nfrm, C.c_float(thop), maxnhar, maxnhar_e, npsd, nch, cfa.ctypes.data_as(C.c_void_p), hm,
                              t, None, 0)
    return t[0], t[1], nfrm * len(idxs)


def cpu_baseline_single(distinct, conf, nutt=16):
    """The reference chain (oracle/_ref, the reference's Release flags) on ONE thread, bounded sample."""
    ct = _conf_tuple(conf)
    nutt = min(nutt, len(distinct["f0"]))
    _G["x"] = [_ref_synth_wave(distinct, ct, b) for b in range(nutt)]
    _G["f0"] = [np.ascontiguousarray(distinct["f0"][b]) for b in range(nutt)]
    _ref_anasynth((ct, [0], 1))                                    # warm-up
    ta, ts, frames = _ref_anasynth((ct, list(range(nutt)), 1))
    return {"value": frames / (ta + ts), "unit": UNIT, "cores": 1, "kind": "reference",
            "analysis_frames_per_s": frames / ta, "synthesis_frames_per_s": frames / ts,
            "sample": "%d utterances x %d frames of the bench workload: llsm_analyze (CZT) + llsm_synthesize of the "
                      "unmodified reference sources + ciglet shim, gcc -Ofast, single thread" % (nutt, conf.nfrm)}


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
    _ref_lib()                                                     # in the parent: see _ref_lib
    nd = 16
    full, conf, distinct = workload(args, distinct=nd)
    ct = _conf_tuple(conf)
    _G["x"] = [_ref_synth_wave(distinct, ct, b) for b in range(nd)]
    _G["f0"] = [np.ascontiguousarray(distinct["f0"][b]) for b in range(nd)]
    ncore = usable_cores()
    per_core = 2
    pool = mp.get_context("fork").Pool(ncore)
    jobs = [(ct, [(c * per_core + i) % nd for i in range(per_core)], 1) for c in range(ncore)]
    step_frames = conf.nfrm * per_core * ncore
    for _ in range(args.warmup):
        pool.map(_ref_anasynth, jobs)
    t0 = time.perf_counter()
    ta = ts = 0.0
    for _ in range(args.steps):
        for a, s, _n in pool.map(_ref_anasynth, jobs):
            ta += a; ts += s
    dt = time.perf_counter() - t0
    pool.close()
    v = step_frames * args.steps / dt
    sample = ("each step = %d utterances x %d frames (%d per host core) of the bench workload through llsm_analyze (CZT) "
              "+ llsm_synthesize of the unmodified reference sources + ciglet shim (gcc -Ofast), one process per core; "
              "%.0f %% of the CPU time is llsm_analyze" % (per_core * ncore, conf.nfrm, per_core, 100 * ta / (ta + ts)))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(args, conf),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncore, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def config_dict(args, conf):
    return {"workload": "BASELINE configs[1] shapes: batch=%d synthetic 2 s utterances per GPU (%d distinct), 44.1 kHz, "
                        "5 ms hop (%d frames), %d harmonics; step = llsm_analyze (f0 refine, CZT harmonics, residual, "
                        "noise PSD, sub-band envelopes) of every waveform + llsm_synthesize (y_sin + y_noise + y) of the "
                        "chunk it produced" % (args.batch, min(args.distinct, args.batch), conf.nfrm, args.nhar),
            "batch_per_gpu": args.batch, "nfrm": conf.nfrm, "nhar": args.nhar, "npsd": conf.npsd,
            "nchannel": conf.nchannel, "fs": conf.fs, "thop": conf.thop,
            "analysis_input": "waveforms synthesised from the synthetic frames (speech-like: harmonics + shaped noise)",
            "noise": "device Philox templates drawn inside the timed step",
            "l2": "inputs + outputs + scratch per step (> 5 GB) exceed the 126 MB L2; no explicit flush",
            "parallelism": ("utterance shards, no data-path collective" if args.shard == "utterances" else
                            "frame-range shards of every utterance + one NCCL all-gather of the overlap-add halos")}


# ------------------------------------------------------------------ algorithmic bytes per kernel
def kernel_bytes(conf, nx):
    """Interface bytes of every kernel of the step per utterance-batch launch (f32, each array once):
    name -> bytes per frame (SURVEY.md 8(d) for the three it lists; the rest from the arrays a kernel must read / write)."""
    hop = nx / float(conf.nfrm + 1) * 4.0                    # bytes of waveform per frame
    nspec = 513.0
    nch, ne = conf.nchannel, conf.maxnhar_e
    hm = 2.0 * conf.maxnhar * 4 + 8                          # ampl, phse, f0, nhar
    return {
        "refine_f0": hop + 8, "harmonic_czt": hop + 4 + hm, "harmonic_pp": hop + 4 + hm,
        "residual_bank": hm + 2 * hop,                       # parameters in, x in, x_res out (the subtraction is fused)
        "noise_spec": 2 * hop + 2 * nspec * 4, "noise_kalman": 4 * nspec * 4,
        "noise_psd_out": 2 * nspec * 4 + 2 * conf.npsd * 4, "subband_iir": 2 * hop + nch * hop,
        "envelope_harmonics": nch * hop + nch * (8 + 8 * ne), "frame_dc": nch * hop + nch * 4,
        "hm_bank": BANK_BYTES_PER_FRAME, "noise_shape": 3812.0, "noise_excitation": 1050.0,
        "noise_shape_fused": 3812.0 + 168.0,                 # excitation evaluated inside the shaper
        "white_fill": nch * 20128 * 4.0 / conf.nfrm, "iir_filtfilt": 2 * nch * 20128 * 4.0 / conf.nfrm,
    }


# ------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import libllsm2_b200 as L
    from libllsm2_b200 import abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU path)")
    try:
        from libllsm2_b200.parallel import bind_to_gpu_numa_node
        bind_to_gpu_numa_node(local)
    except Exception:
        pass
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.shard == "frames":
        return run_b200_frame_shards(args, world, rank, local)
    dev = torch.device("cuda", local)
    full, conf, distinct = workload(args, seed=rank)
    ctx = L.Context(local)
    d = {k: (torch.from_numpy(v).to(dev) if v is not None else None) for k, v in full.items()}
    ny = L.output_length(conf.nfrm, conf.thop, conf.fs)
    frames_per_step = conf.nutt * conf.nfrm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # analysis input: the waveforms of the synthetic frames (untimed)
    x = L.synthesize_l0(ctx, conf, d, white=None, seed=7)["y"].clone()
    f0_in = d["f0"].clone()
    chunk = L.analyze_l0(ctx, conf, x, f0_in)                         # allocates the chunk arrays once
    chunk["nfrm_utt"] = None
    out = {k: torch.empty((conf.nutt, ny), dtype=torch.float32, device=dev) for k in ("y", "y_sin", "y_noise")}
    torch.cuda.synchronize()

    def step(i):
        L.analyze_l0(ctx, conf, x, f0_in, out=chunk)
        L.synthesize_l0(ctx, conf, chunk, white=None, seed=1000 + i, out=out)

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

    # every kernel of the step, live: CUDA events recorded by the library between its launches (a separate pass, so
    # that the timed region above stays free of them)
    kms = {}
    for i in range(max(3, min(args.steps, 7))):
        ctx.set_kernel_timing(True)
        step(10_000 + i)
        for k, v in ctx.kernel_times():
            kms.setdefault(k, []).append(v)
    ctx.set_kernel_timing(False)
    kms = {k: float(np.median(v)) for k, v in kms.items()}

    # end to end through the host-buffer C ABI: pinned waveforms + f0 in, y out, the chunk stays in HBM
    e2e = None
    if not args.no_e2e:
        xh = x.cpu().pin_memory(); fh = f0_in.cpu().pin_memory()
        yh = torch.empty((conf.nutt, ny), dtype=torch.float32).pin_memory()
        h2d = xh.numel() * 4 + fh.numel() * 4
        d2h = yh.numel() * 4
        for i in range(2):
            L.anasynth_host(ctx, conf, xh, fh, seed=i, out={"y": yh})
        barrier()
        t0 = time.perf_counter()
        ne = max(2, min(args.steps, 5))
        for i in range(ne):
            L.anasynth_host(ctx, conf, xh, fh, seed=i, out={"y": yh})
        barrier()
        dt = time.perf_counter() - t0
        e2e_t = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e = {"value": frames_per_step * world * ne / float(e2e_t.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": ne,
               "note": "llsm_b200_anasynth_host: pinned host waveforms + f0 in, analyse, synthesise, y out to pinned "
                       "host memory; the chunk (frame arrays) never leaves the device"}

    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0].item())

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        extras = extra_legs(args, ctx, conf, d, x, f0_in, chunk, out, dev)
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
    kb = kernel_bytes(conf, ny)
    kernels = {}
    for k, v in kms.items():
        bytes_launch = kb.get(k, 0.0) * frames_per_step
        kernels[k] = {"ms": v, "algorithmic_bytes_per_launch": bytes_launch,
                      "achieved_gbs": bytes_launch / (v * 1e-3) / 1e9 if v > 0 else None,
                      "frac": bytes_launch / (v * 1e-3) / 1e9 / peak if v > 0 else None}
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dominant_traffic.json"))).get(dom)
    except Exception:
        pass
    value = frames_per_step * world * args.steps / (ms * 1e-3)
    step_ms = ms / args.steps
    res = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, conf), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                     "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "ms_per_launch": kernels[dom]["ms"],
                     "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_launch"],
                     "whole_step_gbs": (ANA_BYTES_PER_FRAME + SYN_BYTES_PER_FRAME) * frames_per_step / (step_ms * 1e-3) / 1e9,
                     "whole_step_frac": (ANA_BYTES_PER_FRAME + SYN_BYTES_PER_FRAME) * frames_per_step / (step_ms * 1e-3) / 1e9 / peak,
                     "harmonic_bank_frac": kernels.get("hm_bank", {}).get("frac")},
        # the same accounting for every kernel of the step (CUDA events inside the library); "dominant" is the longest
        "kernels": kernels, "dominant_kernel": dom,
    }
    res.update(extras)
    if not args.no_cpu_baseline and world == 1:
        try:
            res["cpu_baseline"] = cpu_baseline_single(distinct, conf)
        except Exception as e:  # the oracle .so travels with the repo; report rather than die
            res["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference",
                                   "sample": "unavailable: %s" % e}
    print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def _dev_time(torch, fn, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def extra_legs(args, ctx, conf, d, x, f0_in, chunk, out, dev):
    """Secondary numbers (rank 0, one GPU): each leg of the step alone, and BASELINE configs[0] (arctic)."""
    import torch
    import libllsm2_b200 as L
    from libllsm2_b200 import abi
    n = max(2, min(args.steps, 5))
    frames = conf.nutt * conf.nfrm
    res = {}
    ana_ms = _dev_time(torch, lambda i: L.analyze_l0(ctx, conf, x, f0_in, out=chunk), n)
    syn_ms = _dev_time(torch, lambda i: L.synthesize_l0(ctx, conf, d, white=None, seed=i, out=out), n)
    res["legs"] = {"analysis": {"ms": ana_ms, "frames_per_s": frames / (ana_ms * 1e-3),
                                "hbm_frac": ANA_BYTES_PER_FRAME * frames / (ana_ms * 1e-3) / 1e9},
                   "synthesis": {"ms": syn_ms, "frames_per_s": frames / (syn_ms * 1e-3),
                                 "hbm_frac": SYN_BYTES_PER_FRAME * frames / (syn_ms * 1e-3) / 1e9}}
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        peak = 6650.0
    for leg in res["legs"].values():
        leg["hbm_frac"] /= peak
    try:
        res["c1"] = c1_leg(args, ctx, dev)
    except Exception as e:
        res["c1"] = {"unavailable": str(e)}
    try:
        res["c3"] = c3_leg(args, ctx, dev)
    except Exception as e:
        res["c3"] = {"unavailable": str(e)}
    return res


def c3_leg(args, ctx, dev):
    """BASELINE configs[2]: batch of 1024 pulse-by-pulse streams through the streaming synthesizer (llsmrt, use_l1), 256
    harmonics, with the growl llsm_pbpeffect (tools/growl_hook.c, the modifier of test/test-pbpeffects.c:70-85, a host C
    function called once per glottal pulse in time order) and without it. Host buffers in and out on both arms
    (llsm_b200_rt_feed_l1_host); with the hook the sequential pulse tracker runs on the host, every sample is still
    produced on the device."""
    import subprocess
    import torch
    import libllsm2_b200 as L
    from libllsm2_b200.synthetic import synth_frames
    so = os.path.join(ROOT, "build", "growl_hook.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "growl_hook.c"), "-lm"])
    hooklib = C.CDLL(so)
    S_, K, NFEED = 1024, 8, 6
    fr, conf = synth_frames(16, K * NFEED, seed=5, nhar=256, maxnhar=256, f0_lo=60, f0_hi=86)
    rep = (S_ + 15) // 16
    frb = {k: (np.ascontiguousarray(np.concatenate([v] * rep, 0)[:S_]) if v is not None else None) for k, v in fr.items()}
    conf.nutt = S_
    d = {k: (torch.from_numpy(v).to(dev) if v is not None and k != "nfrm_utt" else None) for k, v in frb.items()}
    l1d = L.tolayer1(ctx, conf, d, 2048)
    l1 = {k: v.cpu().numpy() for k, v in l1d.items()}
    pbp = np.ones((S_, K * NFEED), np.int32)
    res = {"config": "%d streams x %d frames per feed, 256 harmonics, f0 60-86 Hz, every voiced frame pulse-by-pulse, "
                     "host buffers in / out" % (S_, K)}
    for name, hooked in (("no_effect_device_tracker", False), ("growl_effect_host_tracker", True)):
        conf_k = L.abi.make_conf(S_, K, conf.maxnhar, conf.maxnhar_e, conf.npsd, conf.nchannel, conf.fs, conf.thop)
        rt = L.RtSynth(ctx, conf_k, seed=5, nspec=1025, host_tracker=hooked)
        state = np.zeros((S_, 3), np.uint32)
        state[:, 2] = np.arange(S_, dtype=np.uint32) * np.uint32(2654435761) + np.uint32(12345)
        hook = C.cast(hooklib.growl_hook, C.c_void_p) if hooked else None
        user = state.ctypes.data_as(C.c_void_p) if hooked else None
        ts = []
        for q in range(NFEED):
            sl = slice(q * K, (q + 1) * K)
            fq = {k: (np.ascontiguousarray(v[:, sl]) if v is not None and k != "nfrm_utt" else None) for k, v in frb.items()}
            fq["nhar"] = fq["ampl"] = fq["phse"] = None                 # harmonic model derived from layer 1
            lq = {k: np.ascontiguousarray(v[:, sl]) for k, v in l1.items()}
            t0 = time.perf_counter()
            rt.feed(fq, K, layer1=lq, pbpsyn=np.ascontiguousarray(pbp[:, sl]), hook=hook, user=user)
            ts.append(time.perf_counter() - t0)
        dt = float(np.median(ts[1:]))
        res[name] = {"ms_per_feed": dt * 1e3, "frames_per_s": S_ * K / dt,
                     "x_realtime_per_stream": K * float(conf.thop) / dt}
        if hooked:
            res[name]["pulses_seen_by_hook"] = int(state[:, 0].sum())
        rt.close()
    return res


def c1_leg(args, ctx, dev):
    """BASELINE configs[0]: test/arctic_a0001.wav (committed copy + harness F0: tests/golden/speech_fixtures.npz) through
    analysis + synthesis with the options of test/test-layer0-anasynth.c:29-37, one utterance (latency-bound) and
    64 copies, next to the reference's single-thread C path on the same input."""
    import torch
    import libllsm2_b200 as L
    from libllsm2_b200 import abi
    fx = np.load(os.path.join(ROOT, "tests", "golden", "speech_fixtures.npz"))
    xs = fx["arctic_x"].astype(np.float32) / np.float32(32768.0)
    f0 = fx["arctic_f0"].astype(np.float32)
    fs, nhop = float(fx["arctic_fs"]), int(fx["nhop"])
    F, nx = len(f0), len(xs)
    thop = float(np.float32(nhop) / np.float32(fs))
    res = {"input": "arctic_a0001.wav, %d samples, %d frames (hop %d), maxnhar 400, maxnhar_e 5, npsd 128" % (nx, F, nhop)}
    for method, hm in (("pp", 0), ("czt", 1)):
        r = {}
        for B in (1, 64):
            conf = abi.make_conf(B, F, 400, 5, 128, 4, fs, thop)
            ny = L.output_length(F, thop, fs)
            xh = torch.from_numpy(np.ascontiguousarray(np.tile(xs, (B, 1)))).pin_memory()
            fh = torch.from_numpy(np.ascontiguousarray(np.tile(f0, (B, 1)))).pin_memory()
            yh = torch.empty((B, ny), dtype=torch.float32).pin_memory()
            opt = {"hm_method": hm}
            for i in range(2):
                L.anasynth_host(ctx, conf, xh, fh, options=opt, seed=i, out={"y": yh})
            ts = []
            for i in range(5):
                t0 = time.perf_counter()
                L.anasynth_host(ctx, conf, xh, fh, options=opt, seed=i, out={"y": yh})
                ts.append(time.perf_counter() - t0)
            dt = float(np.median(ts))
            r["batch_%d" % B] = {"e2e_ms": dt * 1e3, "e2e_frames_per_s": B * F / dt,
                                 "x_realtime": B * (nx / fs) / dt}
        # reference, one thread (the path test-layer0-anasynth.c times)
        try:
            lib = _ref_lib()
            cf = np.array([2000.0, 4000.0, 8000.0], np.float32)
            t = (C.c_double * 2)(0.0, 0.0)
            lib.ref_time_anasynth(1, xs.ctypes.data_as(C.c_void_p), nx, C.c_float(fs), f0.ctypes.data_as(C.c_void_p), F,
                                  C.c_float(thop), 400, 5, 128, 4, cf.ctypes.data_as(C.c_void_p), hm, t, None, 0)
            t = (C.c_double * 2)(0.0, 0.0)
            lib.ref_time_anasynth(2, xs.ctypes.data_as(C.c_void_p), nx, C.c_float(fs), f0.ctypes.data_as(C.c_void_p), F,
                                  C.c_float(thop), 400, 5, 128, 4, cf.ctypes.data_as(C.c_void_p), hm, t, None, 0)
            tt = (t[0] + t[1]) / 2
            r["reference_single_thread"] = {"ms": tt * 1e3, "frames_per_s": F / tt, "analysis_ms": t[0] / 2 * 1e3,
                                            "synthesis_ms": t[1] / 2 * 1e3}
            for B in (1, 64):
                r["batch_%d" % B]["speedup_vs_reference_single_thread"] = r["batch_%d" % B]["e2e_frames_per_s"] / (F / tt)
        except Exception as e:
            r["reference_single_thread"] = {"unavailable": str(e)}
        res[method] = r
    return res


def run_b200_frame_shards(args, world, rank, local):
    """north_star's multi-GPU split: every utterance's frame range is cut into `world` shards, each rank synthesises
    its frames (llsm_b200_synthesize_l0_shard) and the ranks exchange the overlap-add halos with ONE NCCL all-gather
    (llsm_b200_halo_exchange: pack kernel -> ncclAllGather -> edge-add kernel). Weak scaling along the frame axis:
    every rank owns --nfrm frames of every utterance, so the utterances are world * nfrm frames long."""
    import torch
    import torch.distributed as dist
    import libllsm2_b200 as L
    from libllsm2_b200 import parallel
    dev = torch.device("cuda", local)
    args_long = argparse.Namespace(**vars(args))
    args_long.nfrm = args.nfrm * world
    args_long.batch = args.batch
    full, conf, _ = workload(args_long, distinct=min(args.distinct, 64), seed=0)    # same utterances on every rank
    ctx = L.Context(local)
    F = conf.nfrm
    lo, hi = rank * args.nfrm, (rank + 1) * args.nfrm
    d = {k: (torch.from_numpy(v).to(dev) if v is not None else None) for k, v in full.items()}
    ny = L.output_length(F, conf.thop, conf.fs)
    out = {k: torch.empty((conf.nutt, ny), dtype=torch.float32, device=dev) for k in ("y", "y_sin", "y_noise")}
    ex = parallel.HaloExchange(ctx, conf, world, rank, out["y"].shape[1]) if world > 1 else None

    def step(i):
        L.synthesize_l0_shard(ctx, conf, d, lo, hi, white=None, seed=1000 + i, out=out)
        if ex is not None:
            ex.exchange(out, lo, hi)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches - l0
    if rank == 0:
        frames = conf.nutt * args.nfrm * world
        cfgd = config_dict(args, conf)
        cfgd["workload"] = ("frame-range sharding: batch=%d utterances of %d frames (%d per GPU), 128 harmonics, layer-0 HM "
                            "synthesis; every rank synthesises its frame range of every utterance, then one NCCL "
                            "all-gather of the overlap-add halos" % (conf.nutt, F, args.nfrm))
        print(json.dumps({
            "metric": "layer0 synthesis frames/sec, frame-range shards + halo all-gather", "value": frames * args.steps / (ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfgd, "clocks": clocks, "gpu_launches": int(launches),
            "halo": {"samples_per_side": ex.halo if ex else 0, "bytes_per_rank": ex.bytes_per_rank if ex else 0,
                     "collective": "ncclAllGather" if ex else None}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
