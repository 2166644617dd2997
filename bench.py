#!/usr/bin/env python3
"""bench.py -- throughput of the H.264 macroblock reconstruction hot path on B200.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
L independent synthetic 1080p P-frame streams ("lanes") per GPU -- random quarter-pel MVs over
all partition shapes, ~25 % coded 4x4 blocks, slice-QP sweep 20..40, deblocking on, one
reference.  One *step* reconstructs one picture of every lane (L pictures): recon_inter (MC +
dequant + IDCT + add) -> deblock -> border, exactly the work of p264_slice_decode steps [3]-[4]
(decoder/decoder.c:623-661) for those pictures.

  value  : pictures/s with the FrameSyntax of every timed step already resident in HBM
  e2e    : pictures/s through the C-ABI with HOST buffers: pinned H2D of every lane's syntax and
           D2H of every reconstructed picture inside the timed region
  roofline / cpu_baseline : see DESIGN.md "Measurement"

`--impl reference` times the reference's own CPU reconstruction (oracle/_ref, the unmodified
sources driven through oracle/ref_harness.c) on all host cores for the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SIZES = {"1080p": (120, 68), "4k": (240, 135), "cif": (22, 18)}
# algorithmic bytes per macroblock, SURVEY.md 8(d): MC+IDCT stage 868 B + 2 B per coefficient slot,
# deblock stage 854 B, whole dense P pipeline 2490 B
BYTES_MC_FIXED, BYTES_DEBLOCK = 868, 854


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", default="1080p", choices=list(SIZES))
    ap.add_argument("--lanes", type=int, default=256, help="independent streams per GPU per launch (the deblock wavefront ramp amortises with lanes: 64 -> 69 k, 256 -> 87 k pictures/s)")
    ap.add_argument("--staged", type=int, default=2, help="distinct pre-staged pictures per lane (cycled)")
    ap.add_argument("--refs", type=int, default=1)
    ap.add_argument("--intra-pct", type=int, default=0, help="side workload: %% of intra macroblocks inside the P pictures (0 = the headline workload)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-parts", default="hcd", help="debug: which legs the e2e step runs (h = H2D, c = compute, d = D2H)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=0, help="pictures per core in the CPU baseline (0 = auto)")
    return ap.parse_args()


def synth_kwargs(args, lane, rank):
    return dict(n_refs=args.refs, seed=stream_seed(rank, lane), first_intra=0, confine_mv=1, intra_pct=args.intra_pct,
                coded_pct=25, max_level=8, mv_range=16, sub8x8=1, skip_pct=5, qp_min=20, qp_max=40, qp_step=2)


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed regions.  NVML is polled every ~2 ms from a thread
    (the device-resident leg lasts tens of milliseconds, too short for `nvidia-smi -lms`); `nvidia-smi` is the
    fallback when the NVML binding is unusable."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.samples, self.proc, self.nvml, self.mx, self.run = [], None, None, None, True
        try:
            import pynvml as N
            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = N.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self.nvml = N
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        N = self.nvml
        while self.run:
            try:
                sm = float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM))
                try:
                    bits = int(N.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.time(), sm, bits))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            r = line.strip().split(", ")
            try:
                bits = 0
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                    if v.strip().lower().startswith("active"):
                        bits |= self.BITS[name]
                self.mx = float(r[1])
                self.samples.append((time.time(), float(r[0]), bits))
            except Exception:
                pass

    def stop(self, windows):
        """windows: [(t0, t1), ...] wall-clock spans of the timed regions"""
        self.run = False
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        if not self.nvml and not self.proc:
            return None
        inside = [(sm, b) for t, sm, b in self.samples if any(t0 - 0.005 <= t <= t1 + 0.005 for t0, t1 in windows)]
        rows = inside or [(sm, b) for _, sm, b in self.samples[-3:]]
        if not rows:
            return None
        bits = 0
        for _, b in rows:
            bits |= b
        return {"sm_mhz": float(np.median([sm for sm, _ in rows])), "sm_max_mhz": self.mx,
                "reasons": sorted(n for n, m in self.BITS.items() if bits & m), "samples": len(rows),
                "source": "nvml" if self.nvml else "nvidia-smi", "in_timed_region": bool(inside)}


# ------------------------------------------------------------------ CPU reference arm
def _cpu_worker(args_tuple):
    """One host core: generate `n` pictures of one stream (untimed), reconstruct them with the
    reference's own functions (oracle/_ref) or the oracle port, return seconds."""
    mb_w, mb_h, refs, seed, n, kind = args_tuple
    sys.path.insert(0, str(ROOT / "tests"))
    import p264decoder_b200 as P
    import _oracle as O

    kw = dict(n_refs=refs, seed=seed, first_intra=0, confine_mv=1, intra_pct=0, coded_pct=25, max_level=8, mv_range=16,
              sub8x8=1, skip_pct=5, qp_min=20, qp_max=40, qp_step=2)
    syn = P.Synth(mb_w, mb_h, **kw)
    frames = [syn.next() for _ in range(n)]
    n_slots = refs + 1
    if kind == "reference":
        eng = O.RefFeed(mb_w, mb_h, n_slots)
    else:
        eng = O.OracleFrames(mb_w, mb_h, n_slots)
    for s in range(n_slots):
        eng.set(s, *P.smooth_picture(16 * mb_w, 16 * mb_h, seed=s))
    fss = [f.syntax() for f in frames]
    t0 = time.perf_counter()
    if kind == "reference":
        for fs in fss:
            O.ref().ref_feed_frame(eng.h, C.byref(fs.hdr), fs.mbs, fs.coefs, 1, 1)
    else:
        for fs in fss:
            O.oracle().orc_recon_frame_flat(C.byref(fs.hdr), fs.mbs, fs.coefs, eng.ptrs, eng.n, 1)
    return time.perf_counter() - t0


def cpu_reference(args, frames_per_core, cores):
    import multiprocessing as mp

    sys.path.insert(0, str(ROOT / "tests"))
    import _oracle as O

    kind = "reference" if O.have_ref() else "port"
    mb_w, mb_h = SIZES[args.size]
    jobs = [(mb_w, mb_h, args.refs, 9000 + i, frames_per_core, kind) for i in range(cores)]
    ctx = mp.get_context("fork")
    if cores == 1:
        times = [_cpu_worker(jobs[0])]
    else:
        with ctx.Pool(cores) as pool:
            times = pool.map(_cpu_worker, jobs)
    fps = cores * frames_per_core / max(times)
    return fps, kind, max(times)


def auto_cpu_frames(size):
    # ~17 pictures/s/core at 1080p for the reference: aim at 10-20 s per core
    return {"1080p": 400, "4k": 100, "cif": 6000}[size]


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.cpu_frames or max(8, auto_cpu_frames(args.size) // 4)
    vals = []
    for i in range(args.warmup + args.steps):
        fps, kind, secs = cpu_reference(args, n, cores)
        if i >= args.warmup:
            vals.append((fps, secs))
    fps = float(np.mean([v[0] for v in vals]))
    unit = f"{args.size}_frames/s"
    sample = f"{cores} worker processes x {n} pictures of the {args.size} synthetic P stream per step, deblock + border + half-pel planes included"
    line = {
        "impl": "reference", "metric": f"reconstructed_{args.size}_frames_per_s", "value": fps, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * float(np.mean([v[1] for v in vals])),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, lanes=args.lanes),
        "cpu_baseline": {"value": fps, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def stream_seed(rank, lane):
    """Streams are disjoint across ranks and lanes: global stream id = rank * 1000 + lane (replicas only)."""
    return 264 + 1000 * rank + lane


def reduce_max(values, dist, device=None):
    """max over ranks of a list of floats (timing is the slowest rank's); identity without a process group"""
    if dist is None:
        return list(values)
    import torch

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def aggregate_value(world, lanes, ms_step):
    """whole-job pictures/s: every rank reconstructs `lanes` pictures per step (weak scaling)"""
    return world * lanes / (ms_step / 1000.0)


def workload_config(args, lanes):
    mb_w, mb_h = SIZES[args.size]
    return {
        "workload": f"BASELINE.json configs[2]: synthetic {args.size} P-frame streams ({16*mb_w}x{16*mb_h} coded), "
                    f"random qpel MVs over all partition shapes incl. sub-8x8, 25% coded 4x4 blocks, slice QP sweep 20..40, "
                    f"deblocking on, {args.refs} reference frame(s)" + (f", {args.intra_pct}% intra macroblocks" if args.intra_pct else ""),
        "lanes_per_gpu": lanes, "staged_pictures_per_lane": args.staged, "mb_per_picture": mb_w * mb_h,
        "l2_policy": "inputs larger than L2: every step streams lanes x (syntax + reference + output picture) >> 126 MB",
        "parallelism": "independent streams (replicas), no collective",
    }


# ------------------------------------------------------------------ GPU arm
def pinned_array(lib, nbytes, dtype):
    p = lib.p264b200_host_alloc(nbytes)
    if not p:
        raise RuntimeError("pinned allocation failed")
    buf = (C.c_uint8 * nbytes).from_address(p)
    return np.frombuffer(buf, dtype=dtype), p


def run_b200(args, rank, world, local_rank):
    import p264decoder_b200 as P

    lib = P.load_library()
    if lib.p264b200_device_count() <= local_rank:
        raise SystemExit("bench.py: no CUDA device -- the reconstruction engine has no CPU fallback")
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_

        torch.cuda.set_device(local_rank)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_
    mb_w, mb_h = SIZES[args.size]
    n_mb = mb_w * mb_h
    L, T = args.lanes, args.staged
    if T % (args.refs + 1):
        T += (args.refs + 1) - T % (args.refs + 1)  # cycling the staged pictures must keep the ring consistent
    n_slots = args.refs + 1

    # ---- generate the workload on the host (pinned, so the e2e leg copies straight from it)
    syns = [P.Synth(mb_w, mb_h, **synth_kwargs(args, l, rank)) for l in range(L)]
    mbs_host, _p1 = pinned_array(lib, T * L * n_mb * 96, np.uint8)
    staged = [[None] * L for _ in range(T)]
    coef_tmp, max_coef, total_coef = {}, 0, 0
    for t in range(T):
        for l in range(L):
            fr = syns[l].next()
            coef_tmp[(t, l)] = fr.coefs
            max_coef = max(max_coef, len(fr.coefs))
            total_coef += len(fr.coefs)
            off = (t * L + l) * n_mb * 96
            mbs_host[off : off + n_mb * 96] = fr.mbs.view(np.uint8)
            staged[t][l] = fr.hdr
    coef_cap = (max_coef + 63) & ~63
    coefs_host, _p2 = pinned_array(lib, T * L * coef_cap * 2, np.int16)
    fss = [[None] * L for _ in range(T)]
    for t in range(T):
        for l in range(L):
            c = coef_tmp[(t, l)]
            off = (t * L + l) * coef_cap
            coefs_host[off : off + len(c)] = c
            fs = P.FrameSyntax()
            fs.hdr = staged[t][l]
            fs.mbs = mbs_host.ctypes.data + (t * L + l) * n_mb * 96
            fs.coefs = coefs_host.ctypes.data + off * 2
            fss[t][l] = fs
    del coef_tmp
    coef_slots_per_step = total_coef / T            # int16 slots per step over all lanes
    # the batched stage call copies the records, the coefficient area up to the last lane's end and the descriptors
    h2d_per_step = L * n_mb * 96 + 2 * ((L - 1) * coef_cap + coef_slots_per_step / L) + L * 440

    eng = P.Engine(mb_w, mb_h, n_slots=n_slots, lanes=L, stage_steps=T, coef_capacity=coef_cap, device=local_rank)
    pics = [P.smooth_picture(16 * mb_w, 16 * mb_h, seed=rank * 1000 + i) for i in range(4)]
    for l in range(L):
        for s in range(n_slots):
            eng.upload(l, s, *pics[(l * n_slots + s) % len(pics)])
    for t in range(T):
        for l in range(L):
            eng.stage(t, l, fss[t][l])
    eng.sync()

    def barrier():
        eng.sync()
        if dist:
            dist.barrier()

    # ---- device-resident leg: `value`
    for i in range(args.warmup):
        eng.recon_step(i % T, L)
    barrier()
    eng.profile_enable(True)
    launches0 = eng.launches
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_wall0 = time.time()
    eng.timer_start()
    for i in range(args.steps):
        eng.recon_step((args.warmup + i) % T, L)
    ms = eng.timer_stop()
    t_wall1 = time.time()
    barrier()
    launches = eng.launches - launches0
    prof = eng.profile_read()
    eng.profile_enable(False)
    clock_windows = [(t_wall0, t_wall1)]

    # ---- end-to-end leg through the C-ABI with host buffers
    e2e_ms, d2h_per_step = None, L * (16 * mb_w * 16 * mb_h * 3 // 2)
    if not args.no_e2e:
        W, H = 16 * mb_w, 16 * mb_h
        out_host, _p3 = pinned_array(lib, L * W * H * 3 // 2, np.uint8)
        yo, uo, vo = 0, W * H, W * H + W * H // 4
        fsz = W * H * 3 // 2
        base = out_host.ctypes.data

        fs_arrays, slot_arrays = [], []
        for t in range(T):
            arr = (P.FrameSyntax * L)(*fss[t])
            fs_arrays.append(arr)
            slot_arrays.append((C.c_int32 * L)(*[fss[t][l].hdr.dst_slot for l in range(L)]))

        def e2e_step(i):
            t = i % T
            rc = 0
            if "h" in args.e2e_parts:
                rc |= lib.p264b200_stage_frames(eng._e, t, L, fs_arrays[t])        # H2D of this step's inputs (pinned)
            if "c" in args.e2e_parts:
                rc |= lib.p264b200_recon_step(eng._e, t, L)
            if "d" in args.e2e_parts:
                rc |= lib.p264b200_frames_download(eng._e, L, slot_arrays[t], base, fsz)  # D2H of every picture
            if rc:
                raise RuntimeError(lib.p264b200_last_error().decode())

        for i in range(args.warmup):
            e2e_step(i)
        barrier()
        t_e0 = time.time()
        eng.timer_start()
        n_e2e = args.steps
        for i in range(n_e2e):
            e2e_step(args.warmup + i)
        e2e_total = eng.timer_stop()
        clock_windows.append((t_e0, time.time()))
        barrier()
        e2e_ms = e2e_total / n_e2e
        assert "d" not in args.e2e_parts or int(out_host[:W].astype(np.int64).sum()) > 0

    # ---- max over ranks
    ms_step = ms / args.steps
    ms_step, e2e_max = reduce_max([ms_step, e2e_ms or 0.0], dist, "cuda")
    e2e_ms = e2e_max if e2e_ms is not None else None
    clocks = sampler.stop(clock_windows) if sampler else None
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    value = aggregate_value(world, L, ms_step)
    unit = f"{args.size}_frames/s"
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    # per-kernel algorithmic bytes per launch (one launch = L pictures)
    alg = {
        "recon_inter": L * n_mb * BYTES_MC_FIXED + 2.0 * coef_slots_per_step,
        "deblock": L * n_mb * BYTES_DEBLOCK,
    }
    kernels = {}
    for k, (kms, kn) in prof.items():
        if kn:
            avg = kms / kn
            kernels[k] = {"avg_ms": avg, "launches": int(kn), "share": kms / ms if ms else None}
            if k in alg:
                kernels[k]["alg_bytes_per_launch"] = alg[k]
                kernels[k]["achieved_gbs"] = alg[k] / (avg * 1e-3) / 1e9
    dominant = max((k for k in kernels if k in alg), key=lambda k: kernels[k]["avg_ms"])
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            rec = json.loads(tf.read_text()).get(f"{args.size}_L{L}", {}).get(dominant)
            traffic = rec and rec.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": dominant, "achieved": kernels[dominant]["achieved_gbs"], "peak": peak, "unit": "GB/s",
        "frac": kernels[dominant]["achieved_gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
        "alg_bytes_per_launch": alg[dominant],
        "pipeline_achieved": (alg["recon_inter"] + alg["deblock"]) / (ms_step * 1e-3) / 1e9,
        "pipeline_frac": (alg["recon_inter"] + alg["deblock"]) / (ms_step * 1e-3) / 1e9 / peak,
    }
    line = {
        "metric": f"reconstructed_{args.size}_frames_per_s", "value": value, "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args, L),
        "mpixel_per_s": value * 16 * mb_w * 16 * mb_h / 1e6,
        "roofline": roofline, "kernels": kernels, "gpu_launches": int(launches), "clocks": clocks,
    }
    if e2e_ms is not None:
        line["e2e"] = {"value": aggregate_value(world, L, e2e_ms), "unit": unit, "h2d_bytes_per_step": int(h2d_per_step),
                       "d2h_bytes_per_step": int(d2h_per_step), "ms_per_step": e2e_ms}
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        n = args.cpu_frames or auto_cpu_frames(args.size)
        fps1, kind, secs1 = cpu_reference(args, n, 1)
        line["cpu_baseline"] = {
            "value": fps1, "unit": unit, "cores": 1, "kind": kind,
            "sample": f"{n} pictures of one {args.size} synthetic P stream on 1 host core ({secs1:.1f} s): reference "
                      f"p264_macroblock_decode + deblock + border + half-pel planes driven from the same FrameSyntax",
            "host_cores_available": cores,
        }
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
