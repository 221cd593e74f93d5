#!/usr/bin/env python
"""Benchmark of the correlation hot path (BASELINE.json metric).

Workload (configs[1] of BASELINE.json): RAFT inference at Sintel resolution 436x1024
(padded to 440x1024 -> 55x128 tokens at 1/8 resolution), D = 256, 4 levels, radius 4,
batch 8 pairs per GPU.  One STEP = one pass of the hot path over one batch:
CorrBlock(fmap1, fmap2) [all-pairs volume + pyramid] followed by ``--iters`` (12)
lookups at fresh coordinates, exactly the calls raft.py:105-107,124 makes per forward.
The convolutional encoder / GRU of RAFT are not part of this path.

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # the reference's CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0).  Keys are described in DESIGN.md ("Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LEVELS, RADIUS, DIM = 4, 4, 256
K_CH = LEVELS * (2 * RADIUS + 1) ** 2
METRIC = "RAFT corr-path pairs/s @436x1024 (12 iters)"


def workload_config(args, H, W):
    """The part of ``config`` both arms share (identical keys and values in both JSON lines)."""
    return {"workload": f"CorrBlock build + {args.iters} lookups per pair, {args.height}x{args.width} px "
                        f"-> {H}x{W} tokens, D={DIM}, L={LEVELS}, r={RADIUS}, batch {args.batch}/GPU",
            "volume": "f32", "coords": "grid + N(0,5^2) 1/8-px flow",
            "l2": "inputs larger than L2 (pyramid 2.1 GB/GPU vs 126 MB)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=12, help="GRU iterations = lookups per pair")
    ap.add_argument("--batch", type=int, default=8, help="image pairs per GPU")
    ap.add_argument("--height", type=int, default=436)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--math", default=os.environ.get("FLOWCORR_MATH", "3xbf16"),
                    choices=["fp32", "3xbf16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rows", action="store_true", help="skip the extra per-row timings (backward, on-demand)")
    return ap.parse_args()


def token_grid(h, w):
    ph, pw = (h + 7) // 8 * 8, (w + 7) // 8 * 8          # InputPadder, utils.py:7-16
    return ph // 8, pw // 8


def synth(B, H, W, iters, seed, device="cpu", pin=False):
    """Feature maps ~ N(0, 1.57^2) (random-init fnet statistics) and a
    coordinate sequence grid + N(0, 5^2) 1/8-px flow (the locality-hostile law)."""
    gen = torch.Generator().manual_seed(seed)
    f1 = 1.57 * torch.randn(B, DIM, H, W, generator=gen)
    f2 = 1.57 * torch.randn(B, DIM, H, W, generator=gen)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    grid = torch.stack([xs, ys], 0).float()[None]
    coords = torch.stack([grid + 5.0 * torch.randn(B, 2, H, W, generator=gen) for _ in range(iters)])
    if pin:
        f1, f2, coords = f1.pin_memory(), f2.pin_memory(), coords.pin_memory()
    return f1.to(device), f2.to(device), coords.to(device)


def lookup_bytes(B, H, W, inbounds_elems=None):
    """Algorithmic bytes of ONE lookup launch: discounted footprint read +
    output write + coords, fp32."""
    n = B * H * W
    foot = inbounds_elems if inbounds_elems is not None else LEVELS * (2 * RADIUS + 2) ** 2
    return n * (foot * 4 + K_CH * 4 + 8)


def inbounds_footprint(coords, H, W):
    """Mean number of in-bounds footprint elements per query (all levels) for the actual
    coordinates: the 'discounted' read of SURVEY.md 8d."""
    tot = 0.0
    c = coords.float()
    for l in range(LEVELS):
        Hl, Wl = H >> l, W >> l
        x0 = torch.floor(c[:, :, 0] / 2 ** l) - RADIUS
        y0 = torch.floor(c[:, :, 1] / 2 ** l) - RADIUS
        nx = (torch.clamp(x0 + 2 * RADIUS + 2, max=Wl) - torch.clamp(x0, min=0)).clamp(min=0)
        ny = (torch.clamp(y0 + 2 * RADIUS + 2, max=Hl) - torch.clamp(y0, min=0)).clamp(min=0)
        tot += float((nx * ny).mean())
    return tot


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions by a thread polling
    NVML (~2 ms period; nvidia-smi takes longer to start than a whole timed region)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        import threading
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop, self._on = threading.Event(), threading.Event()
        self._thread, self.err = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES if it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self._nv, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as e:                      # noqa: BLE001
            self.err = f"nvml unavailable: {e}"

    def _run(self):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            if self._on.is_set():
                try:
                    mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    self.samples.append(mhz)
                    for name, bit in self.BAD.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception as e:              # noqa: BLE001
                    self.err = str(e)
            time.sleep(0.002)

    def resume(self):
        self._on.set()

    def pause(self):
        self._on.clear()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        sm = sorted(self.samples)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
               "samples": len(sm), "reasons": sorted(self.reasons)}
        if self.err and not sm:
            out["reasons"] = [self.err]
        return out


def ncu_traffic(kernel):
    """dram__bytes_read + dram__bytes_write per launch from the committed ncu capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py traffic) or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["hbm_gbs"], p["bf16_tflops"], p["bf16_tflops_sustained"], "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------ CPU arms
def reference_block():
    """-> (CorrBlock class, kind, description).  The reference's OWN class
    (/root/reference/pytorch/core/corr.py, installed verbatim into git-ignored baseline/_ref by
    baseline/install_ref.py, which travels to the GPU box) when it is there -> kind "reference";
    otherwise the library-call port oracle/corr_torch.py (bit-exact to it on CPU) -> kind "port"."""
    try:
        from baseline import install_ref
        p = install_ref.install()
        if p:
            if p not in sys.path:
                sys.path.insert(0, p)
            from core.corr import CorrBlock as RefCorrBlock
            return RefCorrBlock, "reference", "baseline/_ref/pytorch/core/corr.py (unmodified reference CorrBlock)"
    except Exception:                               # noqa: BLE001
        pass
    from oracle import corr_torch
    return corr_torch.TorchCorrBlock, "port", "oracle/corr_torch.py (library-call port of corr.py)"


def cpu_path_rate(H, W, iters, batch, budget_s, threads):
    """The reference's CPU CorrBlock path on the host cores: pairs/s over a bounded sample."""
    Block, kind, what = reference_block()
    torch.set_num_threads(threads)
    f1, f2, coords = synth(batch, H, W, iters, seed=0)
    def one():
        blk = Block(f1, f2, LEVELS, RADIUS)
        for t in range(iters):
            out = blk(coords[t])
        return out
    with torch.no_grad():
        one()
        reps, t0 = 0, time.perf_counter()
        while True:
            one(); reps += 1
            el = time.perf_counter() - t0
            if el >= budget_s:
                break
    return batch * reps / el, reps, el, kind, what


def run_reference(args, H, W):
    """--impl reference: the reference's own CPU implementation of the path -- its unmodified
    ``core.corr.CorrBlock`` from baseline/_ref (else the oracle port) -- on all host cores, on a
    bounded sample of the workload per step.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    Block, kind, what = reference_block()
    sample_b = 2
    f1, f2, coords = synth(sample_b, H, W, args.iters, seed=0)
    def step():
        with torch.no_grad():
            blk = Block(f1, f2, LEVELS, RADIUS)
            for t in range(args.iters):
                blk(coords[t])
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    v = sample_b * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": v,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, H, W),
        "reference_arm": {"math": "fp32 (torch CPU ops)", "what": what,
                          "sample": f"{sample_b} pairs per step on the host cores (bounded sample of the "
                                    f"{args.batch}-pair batch), normalised to pairs/s"},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind,
                         "sample": f"{args.steps} steps x {sample_b} pairs, {what}, torch {torch.__version__} CPU ops"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ GPU arm
def model_e2e(args, B, iters, barrier):
    """-> (dict for the JSON line or None, ms per step on this rank)."""
    import flow_supervisor_b200 as fsb
    try:
        from baseline import install_ref
        p = install_ref.install()
        if p is None:
            return {"unavailable": "baseline/_ref not installed"}, 0.0
        if p not in sys.path:
            sys.path.insert(0, p)
        from core.raft import RAFT
        torch.manual_seed(1234)
        model = RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False)).eval().cuda()
        Hp, Wp = (args.height + 7) // 8 * 8, (args.width + 7) // 8 * 8
        gen = torch.Generator().manual_seed(7)
        im1_h = (torch.rand(B, 3, Hp, Wp, generator=gen) * 255.0).pin_memory()
        im2_h = (torch.rand(B, 3, Hp, Wp, generator=gen) * 255.0).pin_memory()
        flow_h = torch.empty(B, 2, Hp, Wp).pin_memory()
        runner = fsb.RaftRunner(model, iters=iters, graph=True)

        def step():
            a = im1_h.cuda(non_blocking=True)
            b = im2_h.cuda(non_blocking=True)
            _, up = runner(a, b)
            flow_h.copy_(up, non_blocking=True)

        for _ in range(3):
            step()
        barrier()
        n = max(3, min(args.steps, 5))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / n
        info = {"unit": "pairs/s", "what": "images (pinned host) -> RaftRunner(reference RAFT, drop-in block, one CUDA graph) "
                                           f"-> flow (pinned host), {iters} iterations, {args.height}x{args.width}, batch {B}/GPU, "
                                           "torch default conv math",
                "h2d_bytes_per_step": 2 * im1_h.numel() * 4, "d2h_bytes_per_step": flow_h.numel() * 4}
        del runner, model
        torch.cuda.empty_cache()
        return info, ms
    except Exception as e:                              # noqa: BLE001
        return {"unavailable": repr(e)}, 0.0


def main():
    args = parse()
    H, W = token_grid(args.height, args.width)
    if args.impl == "reference":
        run_reference(args, H, W)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (GPU arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import flow_supervisor_b200 as fsb
    from flow_supervisor_b200 import ops, _lib
    fsb.CorrBlock.math = args.math
    math_id = {"fp32": _lib.MATH_FP32, "3xbf16": _lib.MATH_TC_3XBF16, "bf16": _lib.MATH_TC_BF16}[args.math]

    B, iters = args.batch, args.iters
    f1h, f2h, ch = synth(B, H, W, iters, seed=rank, pin=True)           # host (pinned)
    f1, f2, coords = f1h.cuda(), f2h.cuda(), ch.cuda()                  # resident copies
    foot = inbounds_footprint(ch, H, W)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident step: value
    # per-kernel CUDA events ride on every EV_STRIDE-th timed step only (an event record between two launches opens a gap
    # of 1-2 us)
    EV_STRIDE = 4
    ev_steps = [s for s in range(args.steps) if s % EV_STRIDE == 0]
    # three events per sampled step: start, after the build, after the last lookup -- the lookup's figure is the average
    # of its 12 back-to-back launches (an event between every two of them added ~3 us to each)
    ev = {s: [torch.cuda.Event(enable_timing=True) for _ in range(3)] for s in ev_steps}

    def step_resident(events=None):
        if events: events[0].record()
        blk = fsb.CorrBlock(f1, f2, LEVELS, RADIUS)
        if events: events[1].record()
        out = None
        for t in range(iters):
            out = blk(coords[t])
        if events: events[2].record()
        return out

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    lib = _lib.load()
    if sampler: sampler.resume()
    launches0 = lib.fc_kernel_launches()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    t_start.record()
    for s in range(args.steps):
        step_resident(ev.get(s))
    t_end.record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = lib.fc_kernel_launches() - launches0
    if sampler: sampler.pause()
    elapsed_ms = t_start.elapsed_time(t_end)
    build_ms = sum(e[0].elapsed_time(e[1]) for e in ev.values()) / len(ev)
    look_ms = sum(e[1].elapsed_time(e[2]) for e in ev.values()) / (len(ev) * iters)

    # ---- end-to-end step through the public API with HOST buffers: e2e
    out_host = torch.empty(iters, B, K_CH, H, W, dtype=torch.float32).pin_memory()
    h2d = f1h.numel() * 4 * 2 + ch.numel() * 4
    d2h = out_host.numel() * 4

    # H2D of step i+1 (copy stream, double-buffered device inputs) overlaps the D2H of step i's
    # lookup outputs (PCIe is full duplex); every step's copies stay inside the timed region.
    copy_stream = torch.cuda.Stream()
    d2h_stream = torch.cuda.Stream()
    dev_in = [(torch.empty_like(f1), torch.empty_like(f2), torch.empty_like(coords)) for _ in range(2)]
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(in_free[slot])               # the step that last read this slot is done
            a, b, c = dev_in[slot]
            a.copy_(f1h, non_blocking=True); b.copy_(f2h, non_blocking=True); c.copy_(ch, non_blocking=True)
            in_ready[slot].record(copy_stream)

    def run_e2e(n):
        cur = torch.cuda.current_stream()
        for sl in range(2):
            in_free[sl].record(cur)
        upload(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                upload(slot ^ 1)
            cur.wait_event(in_ready[slot])
            a, b, c = dev_in[slot]
            blk = fsb.CorrBlock(a, b, LEVELS, RADIUS)
            for t in range(iters):
                out = blk(c[t])
                # the result leaves on its own stream, so lookup t+1 runs while lookup t's 73 MB cross PCIe
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(d2h_stream):
                    d2h_stream.wait_event(done)
                    out_host[t].copy_(out, non_blocking=True)
                out.record_stream(d2h_stream)
            in_free[slot].record(cur)
        cur.wait_stream(d2h_stream)                             # every step's read-back ends inside the timed region

    run_e2e(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(3, min(args.steps, 10))
    if sampler: sampler.resume()
    e0.record()
    run_e2e(n_e2e)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / n_e2e
    if sampler: sampler.pause()

    # ---- model-level end to end (the metric BASELINE.json names): images in pinned host memory -> RaftRunner over
    # the unmodified reference RAFT (baseline/_ref, random init) with the drop-in block, whole forward in one CUDA
    # graph -> full-resolution flow back in pinned host memory.  H2D and D2H inside the timed region, every step.
    e2e_model, model_ms = model_e2e(args, B, iters, barrier)
    clocks = sampler.stop() if sampler else None

    # ---- max over ranks
    t = torch.tensor([elapsed_ms, e2e_ms, build_ms, look_ms, model_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, build_ms, look_ms, model_ms = t.tolist()
    if e2e_model is not None and model_ms > 0:
        e2e_model.update({"value": world * B / (model_ms * 1e-3), "ms_per_step": model_ms})

    if rank == 0:
        hbm, tf_burst, tf_sust, peak_src = peaks()
        ms_step = elapsed_ms / args.steps
        value = world * B * args.steps / (elapsed_ms * 1e-3)
        lb = lookup_bytes(B, H, W, foot)
        look = {"kernel": "lookup_fwd_kernel", "bound": "hbm", "achieved": lb / (look_ms * 1e-3) / 1e9,
                "peak": hbm, "unit": "GB/s", "traffic": ncu_traffic("lookup_fwd_kernel"), "peak_source": peak_src,
                "bytes_per_launch": lb, "bytes_nominal": lookup_bytes(B, H, W), "ms_per_launch": look_ms,
                "share_of_step": iters * look_ms / ms_step}
        look["frac"] = look["achieved"] / hbm
        # what HBM3e gives this kernel's READ pattern (scattered 128-byte runs of 64-byte patches) with unlimited
        # parallelism, measured by tools/probes/gather_pattern.cu on this pool's B200 (profiles/r01h_gather_pattern_probe.jsonl)
        look["pattern_ceiling"] = {"reads_gbs": 4232.0, "source": "profiles/r01h_gather_pattern_probe.jsonl"}
        N = H * W
        flop = 2.0 * B * N * N * DIM
        pyr_bytes, _ = _lib.pyramid_layout(B, H, W, LEVELS, _lib.VOL_F32)
        # build sits on the ridge: bound = the larger of the two floors for the ALGORITHMIC work
        # (SURVEY.md 8d): useful FLOP / tensor peak vs (fmaps read + pyramid written) / HBM peak
        bld_bytes = pyr_bytes + 2 * B * DIM * N * 4
        issued = flop * (3 if math_id == _lib.MATH_TC_3XBF16 else 1)
        t_mma, t_hbm = flop / (tf_sust * 1e12), bld_bytes / (hbm * 1e9)
        bld = {"kernel": "build (gemm + pyramid)", "launches": ["pack_bf16_kernel", "tc_build_kernel"] if math_id else ["simt"],
               "peak_source": peak_src,
               "traffic": ncu_traffic("tc_build_kernel") if math_id else None,
               "ms_per_launch": build_ms, "share_of_step": build_ms / ms_step,
               "bytes_per_launch": bld_bytes, "flop_per_launch": flop, "flop_issued": issued,
               "floor_ms": {"hbm": 1e3 * t_hbm, "tensor_useful": 1e3 * t_mma,
                            "tensor_issued": 1e3 * issued / (tf_sust * 1e12)},
               "hbm": {"achieved": bld_bytes / (build_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s"},
               "tensor": {"achieved": flop / (build_ms * 1e-3) / 1e12, "issued": issued / (build_ms * 1e-3) / 1e12,
                          "peak": tf_sust, "unit": "TFLOP/s"}}
        for k in ("hbm", "tensor"):
            bld[k]["frac"] = bld[k]["achieved"] / bld[k]["peak"]
        bld["tensor"]["frac_issued"] = bld["tensor"]["issued"] / tf_sust
        which = "hbm" if (t_hbm >= t_mma or not math_id) else "tensor"
        bld.update({"bound": which, **bld[which]})
        dominant, other = (look, bld) if look["share_of_step"] >= bld["share_of_step"] else (bld, look)
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, H, W),
            "arm": {"math": args.math, "parallelism": f"batch-sharded x{world}, no collective",
                    "l2": f"inputs larger than L2 (pyramid {pyr_bytes / 1e9:.2f} GB/GPU, L2 126 MB): no flush needed",
                    "kernel_events": f"CUDA events (start, after the build, after the 12th lookup) on every {EV_STRIDE}th timed step; "
                                     "the lookup's duration is the mean of its 12 back-to-back launches"},
            "lookups_per_s": world * B * N * iters * args.steps / (elapsed_ms * 1e-3),
            "wall_s": wall,
            "roofline": dominant, "roofline_other": other,
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_model": e2e_model,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not args.no_rows and world == 1:
            # the other rows of SURVEY.md section 8 (lookup / build backward at config 3, on-demand at
            # config 5 next to the compiled reference kernel when oracle/_ref holds it): auxiliary,
            # measured after the timed regions above (tools/bench_rows.py states the work per row)
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_rows
                # (the compiled reference kernel of oracle/_ref is timed as a comparison column of the on-demand
                # rows, after and outside every timed region of the product path)
                line["rows"] = bench_rows.collect(reps=5, ref_kernel=True)
            except Exception as e:                      # noqa: BLE001
                line["rows"] = {"error": repr(e)}
        if not args.no_cpu_baseline and world == 1:
            # N=1 only: at N>1 the other ranks would spin in the closing barrier on the host cores being timed
            threads = os.cpu_count() or 1
            rate, reps, el, kind, what = cpu_path_rate(H, W, iters, batch=1, budget_s=12.0, threads=threads)
            line["cpu_baseline"] = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": kind,
                                    "sample": f"{reps} x 1 pair (build + {iters} lookups) in {el:.1f} s, "
                                              f"{what}, torch {torch.__version__} CPU ops"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
