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
    ap.add_argument("--staged", type=int, default=4, help="distinct pre-staged pictures per lane (cycled); picture i of a stream has slice QP 20 + 2 * (i mod 11), "
                    "so 4 staged pictures cover QP 20..26 (both dequant branches), 11 the whole 20..40 sweep")
    ap.add_argument("--refs", type=int, default=1)
    ap.add_argument("--intra-pct", type=int, default=0, help="side workload: %% of intra macroblocks inside the P pictures (0 = the headline workload)")
    ap.add_argument("--pics-per-step", type=int, default=24, help="consecutive pictures of every lane per timed step (device-resident leg): "
                    "20 steps x 24 pictures x 256 lanes keep the timed region above one second")
    ap.add_argument("--e2e-pics-per-step", type=int, default=4, help="pictures of every lane per step of the end-to-end leg (PCIe-bound, ~16 ms per picture of 256 lanes)")
    ap.add_argument("--check-lanes", type=int, default=8, help="lanes whose reference rings are byte-compared with the oracle after the timed loops (0 = no parity gate)")
    ap.add_argument("--no-extras", action="store_true", help="skip the 4K / single-lane sub-runs")
    ap.add_argument("--e2e-wire", default="v2", choices=["v1", "v2"], help="host->device syntax format of the end-to-end leg: v2 = compact FrameSyntax "
                    "(0.62 MB per dense 1080p picture, expanded on the device), v1 = 96-byte records + 16 int16 slots per coded block (2.09 MB)")
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
    sys.path.insert(0, str(ROOT / "tests"))
    import _oracle as O

    if O.have_ref():
        O.ref()          # dlopen oracle/_ref/libp264ref.so in this process too (the workers are forked from it)
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
                    f"random qpel MVs over all partition shapes incl. sub-8x8, 25% coded 4x4 blocks, slice QP sweep 20..40 in steps of 2 per picture "
                    f"(the {args.staged} staged pictures per lane cover QP 20..{20 + 2 * (min(args.staged, 11) - 1)}), "
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


def pick_check_lanes(L, n):
    """first / last lane, both sides of a deblock stream-quad boundary, and lanes spread over the ticket waves
    (tickets are row-group-major over the stream quads, so distant lanes are served by different waves of CTAs)"""
    cand = [0, L - 1, 3, 4, L // 2 - 1, L // 4 + 1, 3 * L // 4 + 2, L // 2 + 2, L // 8 + 3, 5 * L // 8 + 1, 7 * L // 8, 3 * L // 8 + 2]
    out = []
    for c in cand:
        c = min(max(c, 0), L - 1)
        if c not in out:
            out.append(c)
    return sorted(out[:max(0, n)])


def _replay_worker(job):
    """Test-infrastructure leg of the bench (the checker, never the thing measured): replays one lane's picture
    sequence with the CPU oracle and returns the ring (all slots, tight I420 bytes) at the requested picture counts."""
    mb_w, mb_h, n_slots, seeds, frames, checkpoints = job
    sys.path.insert(0, str(ROOT / "tests"))
    import p264decoder_b200 as P
    import _oracle as O

    ring = O.OracleFrames(mb_w, mb_h, n_slots)
    for s, seed in enumerate(seeds):
        ring.set(s, *P.smooth_picture(16 * mb_w, 16 * mb_h, seed=seed))
    objs = []
    for hdr_b, mbs_b, coefs_b in frames:
        hdr = P.FrameHdr.from_buffer_copy(hdr_b)
        objs.append(P.Frame(hdr, np.frombuffer(mbs_b, dtype=P.MB_DTYPE).copy(), np.frombuffer(coefs_b, dtype=np.int16).copy()))
    out, done = [], 0
    for cp in checkpoints:
        while done < cp:
            ring.recon(objs[done % len(objs)])
            done += 1
        out.append([b"".join(np.ascontiguousarray(p).tobytes() for p in fr) for fr in ring.frames])
    return out


class Workload:
    """One engine with `lanes` synthetic streams of one size, T pictures per lane pre-staged in HBM and mirrored in
    pinned host memory (the e2e leg copies straight from it)."""

    def __init__(self, args, size, lanes, refs, staged, rank, local_rank, check_lanes):
        import p264decoder_b200 as P

        self.P, self.lib = P, P.load_library()
        lib = self.lib
        self.size, self.L, self.refs, self.rank = size, lanes, refs, rank
        self.mb_w, self.mb_h = SIZES[size]
        mb_w, mb_h, L = self.mb_w, self.mb_h, lanes
        n_mb = self.n_mb = mb_w * mb_h
        self.n_slots = n_slots = refs + 1
        T = staged
        if T % n_slots:
            T += n_slots - T % n_slots          # cycling the staged pictures must keep the ring consistent
        self.T = T
        self.check = pick_check_lanes(L, check_lanes)
        self.check_frames = {l: [] for l in self.check}
        kw = dict(synth_kwargs(args, 0, rank), n_refs=refs)
        syns = [P.Synth(mb_w, mb_h, **dict(kw, seed=stream_seed(rank, l))) for l in range(L)]
        mbs_host, self._p1 = pinned_array(lib, T * L * n_mb * 96, np.uint8)
        hdrs = [[None] * L for _ in range(T)]
        coef_tmp, max_coef, total_coef = {}, 0, 0
        for t in range(T):
            for l in range(L):
                fr = syns[l].next()
                coef_tmp[(t, l)] = fr.coefs
                max_coef = max(max_coef, len(fr.coefs))
                total_coef += len(fr.coefs)
                off = (t * L + l) * n_mb * 96
                mbs_host[off : off + n_mb * 96] = fr.mbs.view(np.uint8)
                hdrs[t][l] = fr.hdr
                if l in self.check_frames:
                    self.check_frames[l].append((bytes(fr.hdr), fr.mbs.tobytes(), fr.coefs.tobytes()))
        self.coef_cap = coef_cap = (max_coef + 63) & ~63
        coefs_host, self._p2 = pinned_array(lib, T * L * coef_cap * 2, np.int16)
        self.fss = fss = [[None] * L for _ in range(T)]
        for t in range(T):
            for l in range(L):
                c = coef_tmp[(t, l)]
                off = (t * L + l) * coef_cap
                coefs_host[off : off + len(c)] = c
                fs = P.FrameSyntax()
                fs.hdr = hdrs[t][l]
                fs.mbs = mbs_host.ctypes.data + (t * L + l) * n_mb * 96
                fs.coefs = coefs_host.ctypes.data + off * 2
                fss[t][l] = fs
        del coef_tmp
        self.coef_slots_per_step = total_coef / T      # int16 slots per picture of every lane
        # the batched stage call copies the records, the coefficient area up to the last lane's end and the descriptors
        self.h2d_per_pic = L * n_mb * 96 + 2 * ((L - 1) * coef_cap + self.coef_slots_per_step / L) + L * 440
        self.d2h_per_pic = L * (16 * mb_w * 16 * mb_h * 3 // 2)
        self.eng = eng = P.Engine(mb_w, mb_h, n_slots=n_slots, lanes=L, stage_steps=T, coef_capacity=coef_cap, device=local_rank)
        self.pic_seeds = [rank * 1000 + i for i in range(4)]
        pics = [P.smooth_picture(16 * mb_w, 16 * mb_h, seed=s) for s in self.pic_seeds]
        for l in range(L):
            for s in range(n_slots):
                eng.upload(l, s, *pics[(l * n_slots + s) % len(pics)])
        for t in range(T):
            for l in range(L):
                eng.stage(t, l, fss[t][l])
        eng.sync()
        self.pic = 0                                     # pictures reconstructed per lane so far
        self.fs_arrays = [(P.FrameSyntax * L)(*fss[t]) for t in range(T)]
        self.slot_arrays = [(C.c_int32 * L)(*[fss[t][l].hdr.dst_slot for l in range(L)]) for t in range(T)]
        self.out_host = None
        self.v2_arrays = None

    def pack_v2(self):
        """the e2e leg's inputs in the compact wire format: every staged picture packed once (what a parser that emits v2
        directly would hand over), the lanes of one step back to back in pinned memory = one H2D copy per step"""
        P, lib, L = self.P, self.lib, self.L
        self.v2_arrays, self.v2_bytes, self._p4 = [], [], []
        for t in range(self.T):
            need = [lib.p264b200_pack_v2_bound(self.mb_w, self.mb_h, self.fss[t][l].hdr.n_coef) for l in range(L)]
            tmp = np.zeros(max(need) + 16, np.uint8)
            o = (-tmp.ctypes.data) % 16
            packed = []
            for l in range(L):
                v2, _ = P.pack_v2(self.fss[t][l], tmp[o : o + need[l]])
                packed.append((v2, tmp[o : o + v2.blob_bytes].copy()))
            total = sum(len(b) for _, b in packed)
            arena, ptr = pinned_array(lib, total + 16, np.uint8)
            self._p4.append(ptr)
            at = (-arena.ctypes.data) % 16
            arr = (P.FrameSyntaxV2 * L)()
            for l, (v2, b) in enumerate(packed):
                arena[at : at + len(b)] = b
                v2.blob = arena.ctypes.data + at
                arr[l] = v2
                at += len(b)
            self.v2_arrays.append(arr)
            self.v2_bytes.append(total + L * (440 + 40))
        self.h2d_per_pic = float(np.mean(self.v2_bytes))

    # one picture of every lane, syntax already resident in HBM
    def recon(self):
        self.eng.recon_step(self.pic % self.T, self.L)
        self.pic += 1

    # one picture of every lane through the C-ABI with host buffers
    def e2e_picture(self, parts="hcd"):
        lib, eng, t, L = self.lib, self.eng, self.pic % self.T, self.L
        if self.out_host is None:
            W, H = 16 * self.mb_w, 16 * self.mb_h
            self.fsz = W * H * 3 // 2
            self.out_host, self._p3 = pinned_array(lib, L * self.fsz, np.uint8)
        rc = 0
        if "h" in parts:                                                            # H2D of this picture's inputs (pinned)
            if self.v2_arrays is not None:
                rc |= lib.p264b200_stage_frames_v2(eng._e, t, L, self.v2_arrays[t])
            else:
                rc |= lib.p264b200_stage_frames(eng._e, t, L, self.fs_arrays[t])
        if "c" in parts:
            rc |= lib.p264b200_recon_step(eng._e, t, L)
        if "d" in parts:
            rc |= lib.p264b200_frames_download(eng._e, L, self.slot_arrays[t], self.out_host.ctypes.data, self.fsz)  # D2H of every picture
        if rc:
            raise RuntimeError(lib.p264b200_last_error().decode())
        self.pic += 1

    def ring_bytes(self, lane):
        return [b"".join(np.ascontiguousarray(p).tobytes() for p in self.eng.download(lane, s)) for s in range(self.n_slots)]

    def replay_jobs(self, lanes, checkpoints):
        n = self.n_slots
        return [(self.mb_w, self.mb_h, n, [self.pic_seeds[(l * n + s) % len(self.pic_seeds)] for s in range(n)], self.check_frames[l], list(checkpoints))
                for l in lanes]

    def close(self):
        self.eng.close()
        for p in [self._p1, self._p2, getattr(self, "_p3", None)] + list(getattr(self, "_p4", [])):
            if p:
                self.lib.p264b200_host_free(p)


def run_replays(jobs):
    import multiprocessing as mp

    if not jobs:
        return []
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        return pool.map(_replay_worker, jobs)


def timed_device_leg(wl, warmup, steps, K):
    """`steps` timed steps of K pictures per lane (syntax resident in HBM); returns (ms, launches, per-kernel profile, window)"""
    eng = wl.eng
    for _ in range(warmup * K):
        wl.recon()
    eng.sync()
    eng.profile_enable(True)
    l0 = eng.launches
    t0 = time.time()
    eng.timer_start()
    for _ in range(steps * K):
        wl.recon()
    ms = eng.timer_stop()
    t1 = time.time()
    eng.sync()
    prof = eng.profile_read()
    eng.profile_enable(False)
    return ms, eng.launches - l0, prof, (t0, t1)


def kernel_table(prof, ms_total, alg):
    kernels = {}
    for k, (kms, kn) in prof.items():
        if kn:
            avg = kms / kn
            kernels[k] = {"avg_ms": avg, "launches": int(kn), "share": kms / ms_total if ms_total else None}
            if k in alg:
                kernels[k]["alg_bytes_per_launch"] = alg[k]
                kernels[k]["achieved_gbs"] = alg[k] / (avg * 1e-3) / 1e9
    return kernels


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
    L, K, Ke = args.lanes, max(1, args.pics_per_step), max(1, args.e2e_pics_per_step)
    wl = Workload(args, args.size, L, args.refs, args.staged, rank, local_rank, args.check_lanes)
    eng = wl.eng

    def barrier():
        eng.sync()
        if dist:
            dist.barrier()

    # ---- device-resident leg: `value`
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches, prof, win = timed_device_leg(wl, args.warmup, args.steps, K)
    barrier()
    clock_windows = [win]
    pics_after_device = wl.pic
    dev_rings = {l: wl.ring_bytes(l) for l in wl.check}          # downloaded after the timed loop, compared below

    # ---- end-to-end leg through the C-ABI with host buffers
    e2e_ms, e2e_lanes, e2e_rings, e2e_last = None, [], {}, {}
    if not args.no_e2e:
        if args.e2e_wire == "v2":
            wl.pack_v2()
        for _ in range(args.warmup * Ke):
            wl.e2e_picture(args.e2e_parts)
        barrier()
        t_e0 = time.time()
        eng.timer_start()
        for _ in range(args.steps * Ke):
            wl.e2e_picture(args.e2e_parts)
        e2e_total = eng.timer_stop()
        clock_windows.append((t_e0, time.time()))
        barrier()
        e2e_ms = e2e_total / args.steps
        assert "d" not in args.e2e_parts or int(wl.out_host[: 16 * mb_w].astype(np.int64).sum()) > 0
        if args.e2e_parts == "hcd":
            e2e_lanes = wl.check[:1] + wl.check[-1:] if len(wl.check) > 1 else wl.check
            e2e_rings = {l: wl.ring_bytes(l) for l in e2e_lanes}
            e2e_last = {l: wl.out_host[l * wl.fsz : (l + 1) * wl.fsz].tobytes() for l in e2e_lanes}   # what the last D2H delivered
    pics_after_e2e = wl.pic

    # ---- max over ranks
    ms_step = ms / args.steps
    ms_step, e2e_max = reduce_max([ms_step, e2e_ms or 0.0], dist, "cuda")
    e2e_ms = e2e_max if e2e_ms is not None else None
    clocks = sampler.stop(clock_windows) if sampler else None

    # ---- parity gate at the timed shape: the oracle replays the same picture sequences (after the timing, on the host cores)
    parity = None
    if wl.check:
        jobs = wl.replay_jobs(wl.check, [pics_after_device] + ([pics_after_e2e] if e2e_lanes else []))
        if e2e_lanes:   # only the e2e lanes need the second checkpoint
            jobs = [j if l in e2e_lanes else j[:5] + ([pics_after_device],) for l, j in zip(wl.check, jobs)]
        t_r0 = time.time()
        res = run_replays(jobs)
        bad = []
        last_slot = wl.fss[(pics_after_e2e - 1) % wl.T][0].hdr.dst_slot
        for l, r in zip(wl.check, res):
            if r[0] != dev_rings[l]:
                bad.append(f"lane {l}: reference ring after {pics_after_device} device-resident pictures differs from the oracle")
            if l in e2e_lanes:
                if r[1] != e2e_rings[l]:
                    bad.append(f"lane {l}: reference ring after the end-to-end leg ({pics_after_e2e} pictures) differs from the oracle")
                if r[1][last_slot] != e2e_last[l]:
                    bad.append(f"lane {l}: last downloaded picture of the end-to-end leg differs from the oracle")
        parity = {"ok": not bad, "lanes_checked": len(wl.check), "lanes": wl.check, "ring_slots_compared": wl.n_slots,
                  "pictures_replayed_per_lane": pics_after_device, "e2e_lanes_checked": len(e2e_lanes),
                  "e2e_pictures_replayed_per_lane": pics_after_e2e if e2e_lanes else 0,
                  "checker": "oracle/liboracle.so (CPU restatement) replaying the timed sequences", "seconds": round(time.time() - t_r0, 1)}
        if bad:
            parity["mismatches"] = bad[:8]
    ok_all = reduce_max([0.0 if (parity is None or parity["ok"]) else 1.0], dist, "cuda")[0] == 0.0
    if parity is not None and not ok_all:
        parity["ok"] = False
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        if not ok_all:
            raise SystemExit(1)
        return

    value = aggregate_value(world, L * K, ms_step)
    unit = f"{args.size}_frames/s"
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    # per-kernel algorithmic bytes per launch (one launch = one picture of every lane = L pictures)
    alg = {
        "recon_inter": L * n_mb * BYTES_MC_FIXED + 2.0 * wl.coef_slots_per_step,
        "deblock": L * n_mb * BYTES_DEBLOCK,
    }
    kernels = kernel_table(prof, ms, alg)
    dominant = max((k for k in kernels if k in alg), key=lambda k: kernels[k]["avg_ms"])
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            rec = json.loads(tf.read_text()).get(f"{args.size}_L{L}", {}).get(dominant)
            traffic = rec and rec.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    ms_pic = ms_step / K                                  # one picture of every lane
    pipe = (alg["recon_inter"] + alg["deblock"]) / (ms_pic * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": dominant, "achieved": kernels[dominant]["achieved_gbs"], "peak": peak, "unit": "GB/s",
        "frac": kernels[dominant]["achieved_gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
        "alg_bytes_per_launch": alg[dominant], "pipeline_achieved": pipe, "pipeline_frac": pipe / peak,
    }
    cfg = workload_config(args, L)
    cfg["pictures_per_lane_per_step"] = K
    line = {
        "metric": f"reconstructed_{args.size}_frames_per_s", "value": value, "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "ms_per_picture_of_all_lanes": ms_pic,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
        "mpixel_per_s": value * 16 * mb_w * 16 * mb_h / 1e6,
        "roofline": roofline, "kernels": kernels, "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
    }
    if e2e_ms is not None:
        line["e2e"] = {"value": aggregate_value(world, L * Ke, e2e_ms), "unit": unit, "h2d_bytes_per_step": int(wl.h2d_per_pic * Ke),
                       "d2h_bytes_per_step": int(wl.d2h_per_pic * Ke), "ms_per_step": e2e_ms, "pictures_per_lane_per_step": Ke,
                       "wire_format": f"FrameSyntax {args.e2e_wire}" + (" (compact: partition vectors + significance masks + int8 levels, expanded on the device)" if args.e2e_wire == "v2" else "")}
    wl.close()
    del wl

    # ---- extras (N=1 only; short sub-runs outside the headline's timed regions): 4K / multi-reference and one-lane latency
    if world == 1 and not args.no_extras and args.size == "1080p" and not args.intra_pct:
        line["extra"] = run_extras(args, rank, local_rank, peak)
        if any(isinstance(v, dict) and v.get("parity") and not v["parity"]["ok"] for v in line["extra"].values()):
            ok_all = False
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
    if not ok_all:
        print("bench.py: PARITY MISMATCH against the oracle -- the numbers above are void", file=sys.stderr)
        raise SystemExit(1)


def run_extras(args, rank, local_rank, peak):
    """configs[3] (4K, 4 references) and the single-lane latency as short sub-runs, each with its own parity check."""
    import copy

    out = {}
    # 4K, 4 references, 48 lanes
    a4 = copy.copy(args)
    a4.size, a4.refs = "4k", 4
    L4, K4, steps4, warm4 = 48, 2, 8, 3
    wl = Workload(a4, "4k", L4, 4, 5, rank, local_rank, 2)
    ms, _, prof, _ = timed_device_leg(wl, warm4, steps4, K4)
    n_mb = wl.n_mb
    alg = {"recon_inter": L4 * n_mb * BYTES_MC_FIXED + 2.0 * wl.coef_slots_per_step, "deblock": L4 * n_mb * BYTES_DEBLOCK}
    kt = kernel_table(prof, ms, alg)
    ms_pic = ms / (steps4 * K4)
    rings = {l: wl.ring_bytes(l) for l in wl.check}
    res = run_replays(wl.replay_jobs(wl.check, [wl.pic]))
    ok = all(r[0] == rings[l] for l, r in zip(wl.check, res))
    pipe = (alg["recon_inter"] + alg["deblock"]) / (ms_pic * 1e-3) / 1e9
    out["4k"] = {"workload": "BASELINE.json configs[3]: synthetic 4K (3840x2160) P streams, 4 reference frames, deblocking on", "lanes": L4,
                 "value": L4 / (ms_pic * 1e-3), "unit": "4k_frames/s", "ms_per_picture_of_all_lanes": ms_pic,
                 "roofline": {"kernel": "recon_inter", "frac": kt["recon_inter"]["achieved_gbs"] / peak, "achieved": kt["recon_inter"]["achieved_gbs"],
                              "deblock_frac": kt["deblock"]["achieved_gbs"] / peak, "pipeline_frac": pipe / peak},
                 "parity": {"ok": ok, "lanes_checked": len(wl.check), "pictures_replayed_per_lane": wl.pic}}
    wl.close()
    del wl
    # one 1080p lane: the wavefront-latency floor of a single stream (hard part 1)
    wl = Workload(args, "1080p", 1, args.refs, 2, rank, local_rank, 1)
    ms, _, prof, _ = timed_device_leg(wl, 3, 40, 1)
    rings = {l: wl.ring_bytes(l) for l in wl.check}
    res = run_replays(wl.replay_jobs(wl.check, [wl.pic]))
    ok = all(r[0] == rings[l] for l, r in zip(wl.check, res))
    out["single_lane"] = {"workload": "one 1080p stream, one picture per launch sequence (syntax resident in HBM)", "ms_per_picture": ms / 40,
                          "value": 1000.0 * 40 / ms, "unit": "1080p_frames/s",
                          "kernels_ms": {k: v[0] / v[1] for k, v in prof.items() if v[1]}, "parity": {"ok": ok, "pictures_replayed_per_lane": wl.pic}}
    wl.close()
    out["real_bitstream"] = real_bitstream_extra()
    return out


def real_bitstream_extra(n_streams=64, n_pic=24):
    """The same kind of pictures as a REAL Annex-B bitstream (csrc/host/writer.cc: Baseline / CAVLC, one reference, partitions >= 8x8),
    decoded from the byte stream by N concurrent streams through the multi-stream decoder: host entropy decode on all cores + batched
    GPU reconstruction.  This path is bound by the host parser; it is reported, not the headline.  Parity: stream 0's YUV against the
    unmodified reference CLI (oracle/_ref/p264dec_ref, the checker) when that binary travelled with the repository."""
    import hashlib
    import tempfile

    sys.path.insert(0, str(ROOT / "tools"))
    import make_stream

    exe = ROOT / "p264decoder_b200" / "lib" / "p264dec_multi"
    ref = ROOT / "oracle" / "_ref" / "p264dec_ref"
    if not exe.exists():
        return {"unavailable": "p264dec_multi not built"}
    with tempfile.TemporaryDirectory() as td:
        src = Path(td) / "synth1080.264"
        nbytes = make_stream.make(src, "1080p", n_pic)
        r = subprocess.run([str(exe), "-n", str(n_streams), str(src)], capture_output=True, text=True)
        steady = [l for l in r.stderr.splitlines() if "after the first step" in l]
        whole = [l for l in r.stderr.splitlines() if "decoding speed" in l]
        if r.returncode or not steady:
            return {"error": r.stderr[-300:]}
        res = {"workload": f"{n_streams} concurrent copies of a written 1080p CAVLC stream ({n_pic} pictures, {nbytes // n_pic} bytes per picture), host parse on "
                           f"{os.cpu_count()} cores + batched GPU reconstruction + download",
               "value": float(steady[0].split(":")[1].split()[0]), "unit": "1080p_frames/s",
               "whole_process_value": float(whole[0].split(":")[1].split()[0]) if whole else None, "bound": "host parser"}
        r2 = subprocess.run([str(exe), "-n", "2", "-o", str(Path(td) / "out"), str(src)], capture_output=True, text=True)
        if r2.returncode == 0 and ref.exists():
            r3 = subprocess.run([str(ref), "-d", str(src), str(Path(td) / "ref.yuv")], capture_output=True, text=True)
            if r3.returncode == 0:
                a = hashlib.md5((Path(td) / "out1.yuv").read_bytes()).hexdigest()
                b = hashlib.md5((Path(td) / "ref.yuv").read_bytes()).hexdigest()
                res["parity"] = {"ok": a == b, "checker": "unmodified reference CLI (oracle/_ref/p264dec_ref), md5 of the whole YUV"}
        return res


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
