#!/usr/bin/env python
"""bench.py -- Mpixel/s of gg's scene-rasterisation hot path on B200 (ggcuda) and on the host CPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" is one pass of the hot path (packed scene already in HBM -> flatten incl. stroke expansion -> bin ->
coarse/PTCL -> fine -> RGBA8 band in HBM, then the assembly of the bands in every rank's frame when N > 1) over the
synthetic scene `config3` (BASELINE.json configs[2]: 3840x2160, 10 000 filled + stroked Bezier paths, 29 blend modes,
layers, clips). For N GPUs the canvas is N copies of that frame stacked vertically (weak scaling: each GPU owns one 4K band
of 135 tile rows); every rank holds the whole encoding and renders only its band. Band assembly: up to 4 GPUs the fine
kernel stores its band into every rank's frame itself (symmetric memory: NVSwitch multicast / peer pointers) and a barrier
follows; at 8 GPUs one NCCL all-gather (GG_BANDS=p2p|p2p_nomc|nccl overrides). `value` is device time (CUDA events, max over
ranks); `e2e` is the same frame through the public host API with host buffers (scene ingest + H2D + pipeline + D2H inside
the timed region). `--impl reference` times the CPU restatement of the reference's pipeline on the host cores.
"""
import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def fine_algorithmic_bytes(ptcl_off, ptcl, w, h):
    """B_fine = 4 W H + 4 sum(ptcl words) + 20 sum over tile-fills of seg_count."""
    words = 0
    segs = 0
    n = len(ptcl_off)
    for t in range(n):
        o = int(ptcl_off[t]) + 1
        words += 1
        while True:
            tag = int(ptcl[o])
            if tag == 0:
                words += 1
                break
            if tag == 1:
                segs += int(ptcl[o + 1]) >> 1
                o += 4; words += 4
            elif tag == 3 or tag == 10:
                o += 1; words += 1
            elif tag == 5:
                o += 2; words += 2
            elif tag == 11:
                o += 3; words += 3
            else:
                break
    return 4 * w * h + 4 * words + 20 * segs, words, segs


def fine_bytes_fast(ptcl_off, ptcl, w, h):
    """Vectorised version of fine_algorithmic_bytes (walks all tiles in lock-step)."""
    pos = ptcl_off.astype(np.int64) + 1
    active = np.ones(len(pos), dtype=bool)
    words = len(pos)
    segs = 0
    size = np.zeros(16, dtype=np.int64)
    size[[0, 1, 3, 5, 10, 11]] = [1, 4, 1, 2, 1, 3]
    while active.any():
        idx = np.nonzero(active)[0]
        tags = ptcl[pos[idx]]
        is_fill = tags == 1
        if is_fill.any():
            segs += int((ptcl[pos[idx[is_fill]] + 1] >> 1).sum())
        sz = size[np.minimum(tags, 15)]
        words += int(sz.sum())
        pos[idx] += sz
        active[idx[(tags == 0) | (sz == 0)]] = False
    return 4 * w * h + 4 * words + 20 * segs, words, segs


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port of internal/gpu/tilecompute; the Go
    original cannot be built here: no Go toolchain), all host threads, on this arm's config."""
    if rank != 0:
        return
    from oracle import twin as T
    from gg_b200 import _lib, scenes
    # Bounded sample: one 3840x2160 band (the per-GPU share of the weak-scaling canvas), whatever N is -- the CPU
    # throughput in Mpix/s does not depend on how many bands the canvas has, and a step stays at a few seconds.
    enc, w, h = scenes.config3(bands=1)
    # scene preparation (not timed): the packed scene both arms consume, from a host-only context
    hc = _lib.Context(-1)
    hc.begin(w, h)
    hc.add_encoding(*enc.streams())
    words, layout = hc.pack_host()
    hc.close()
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        T.render_packed(words, layout, w, h, (0, 0, 0, 0), threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, tm = T.render_packed(words, layout, w, h, (0, 0, 0, 0), threads)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    val = w * h / 1e6 / dt
    line = {"impl": "reference", "metric": "Mpix/s", "value": val, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config3_4k_10k_paths_blend_layers_clips", "width": w, "height": h * max(1, args.gpus),
                       "bands": max(1, args.gpus)},
            "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": threads, "kind": "port",
                             "sample": f"one {w}x{h} band per step (flatten+coarse on 1 thread, fine on {threads} threads); "
                                       "port of internal/gpu/tilecompute (the Go original cannot be built here)",
                             "stage_s": {k: tm[k] for k in ("t_flatten", "t_coarse", "t_fine")}},
            "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-baseline", type=int, default=1, help="time the oracle on the host cores at N=1 (0 to skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # Watchdog: a wedged collective or driver call must not hold a GPU box until the caller's limit; dump every
    # thread's stack and exit instead.
    faulthandler.dump_traceback_later(int(os.environ.get("GG_BENCH_WATCHDOG_S", "900")), exit=True)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from gg_b200 import _lib, bands, scenes
    from gg_b200.accelerator import CUDAAccelerator, GPURenderTarget

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    args.warmup = max(3, args.warmup)

    enc, w, h = scenes.config3(bands=world)
    streams = enc.streams()
    y0, y1 = bands.band_rows(h, world, rank)

    ctx = _lib.Context(local_rank)
    stream = torch.cuda.Stream()          # kernels, events and the all-gather all go through this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_timing(True)
    ctx.begin(w, h)
    ctx.add_encoding(*streams)
    ctx.set_band(y0, y1)
    ctx.upload()
    # Full canvas on every rank. Up to 4 GPUs: a symmetric allocation -- fine stores its band into every rank's frame
    # itself (NVSwitch multicast where offered, else peer stores) and a barrier replaces the all-gather. 8 GPUs: one NCCL
    # all-gather of the bands (GG_BANDS=p2p|p2p_nomc|nccl overrides; NCCL is also the fallback if symmetric memory cannot
    # be set up on this box).
    sym, assemble_kind = None, "single"
    # measured on 8 x B200 (round 1, ms per step, fused stores vs all-gather): N=2 1.62 / 1.67, N=4 1.71 / 1.76,
    # N=8 2.08 / 1.98 -- fine's 64-byte row pieces make poor NVLink packets once seven peers share the switch
    band_mode = os.environ.get("GG_BANDS", "auto")
    if band_mode == "auto":
        band_mode = "p2p" if world <= 4 else "nccl"
    if world > 1 and band_mode != "nccl":
        try:
            sym = bands.SymmetricFrame(w, h, world, rank, f"cuda:{local_rank}")
            if band_mode == "p2p_nomc":
                sym.multicast = False
            assemble_kind = "fine stores into all frames (multimem.st over NVSwitch multicast) + barrier" if sym.multicast else \
                "fine stores into all frames (peer memory over NVLink) + barrier"
        except Exception as e:   # noqa: BLE001
            if rank == 0:
                print(f"symmetric memory unavailable ({type(e).__name__}: {e}); using the NCCL all-gather", file=sys.stderr)
            sym = None
    if world > 1 and band_mode != "nccl":   # every rank must have it, or nobody uses it
        ok = torch.tensor([1 if sym is not None else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            sym = None
    if sym is not None:
        frame, band = sym.frame, sym.band()
    else:
        frame = bands.alloc_frame(w, h, world, "cuda")
        band = bands.band_view(frame, h, world, rank)     # fine writes its band straight into the gather buffer
        if world > 1:
            assemble_kind = "NCCL all_gather_into_tensor"
    stride = w * 4
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    e_mid = torch.cuda.Event(enable_timing=True)

    def step():
        if sym is not None:
            if sym.multicast:
                ctx.render_device_multi(band.data_ptr(), [sym.multicast_band], stride, _lib.KEEP_SCENE, multicast=True)
            else:
                ctx.render_device_multi(band.data_ptr(), sym.peer_bands, stride, _lib.KEEP_SCENE)
            e_mid.record(stream)
            sym.barrier()
        else:
            ctx.render_device(band.data_ptr(), stride, _lib.KEEP_SCENE)
            e_mid.record(stream)
            bands.assemble(frame, h, world, rank)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    st = ctx.stats()
    launches_per_step = st["kernel_launches"]

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms, fine_ms, stage_ms, own_ms = 0.0, 0.0, np.zeros(4), 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.steps):
        flush_buf.fill_(1)          # L2 flush between timed iterations (not timed)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        own_ms += e0.elapsed_time(e_mid)
        s = ctx.stats()
        fine_ms += s["ms_fine"]
        stage_ms += [s["ms_front"], s["ms_binning"], s["ms_coarse"], s["ms_fine"]]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms, own_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    own_max_ms = float(t[1].item()) / args.steps   # slowest rank's own pipeline, before it waits for the others
    t = t[:1]
    ms_per_step = float(t.item()) / args.steps
    value = w * h / 1e6 / (ms_per_step / 1e3)

    # ---- roofline: algorithmic bytes / CUDA-event time, per stage; the headline object is
    #      the stage that takes longest. Fine's bytes count only what lies after each tile's restart point (the
    #      part of the PTCL it has to execute), so skipping dead commands does not inflate its GB/s.
    poff = ctx.debug_read(_lib.BUF_PTCL_OFF, np.uint32)
    ptcl = ctx.debug_read(_lib.BUF_PTCL, np.uint32)
    rst = ctx.debug_read(_lib.BUF_RESTART, np.uint32).reshape(-1, 2)[:, 0]
    lay = ctx.debug_read(_lib.BUF_LAYOUT, _lib.LAYOUT)[0]
    band_h = min(y1 * 16, h) - y0 * 16
    _, words_all, _ = fine_bytes_fast(poff, ptcl, w, band_h)
    b_fine, words, segs = fine_bytes_fast(poff + np.maximum(rst, 1) - 1, ptcl, w, band_h)
    peak, peak_src = _peaks()
    n_pd = int(lay["draw_tag_base"] - lay["path_data_base"])
    n_tr = int(lay["style_base"] - lay["transform_base"]) // 6
    c = {k: int(st[k]) for k in ("n_lines", "n_path_tiles", "n_segments", "n_hits", "n_ptcl_words", "n_draws", "n_tag_bytes")}
    stage_bytes = {
        "front": c["n_tag_bytes"] + 4 * n_pd + 24 * n_tr + 20 * c["n_lines"] + 6 * c["n_tag_bytes"] + 20 * c["n_draws"],
        "binning": 20 * c["n_lines"] + 8 * c["n_segments"] + 8 * c["n_path_tiles"] + 32 * c["n_path_tiles"] + 48 * c["n_segments"],
        "coarse": 16 * c["n_draws"] + 8 * c["n_path_tiles"] + 4 * words_all,
        "fine": b_fine,
    }
    stages = {}
    for k, ms in zip(("front", "binning", "coarse", "fine"), stage_ms / args.steps):
        gbs = stage_bytes[k] / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        stages[k] = {"ms": float(ms), "algorithmic_bytes": int(stage_bytes[k]), "achieved": gbs, "frac": gbs / peak}
    dom = max(stages, key=lambda k: stages[k]["ms"])
    # DRAM bytes of the dominant stage's main kernel from the committed `ncu --set full` capture of this workload
    # (profiles/*_traffic.json; null if that kernel was not captured)
    traffic = None
    try:
        tj = json.load(open(sorted(p for p in (os.path.join(ROOT, "profiles", f) for f in os.listdir(os.path.join(ROOT, "profiles")))
                                   if p.endswith("_traffic.json"))[-1]))
        kname = {"front": "flatten_subdivide_kernel", "binning": "path_count_kernel", "coarse": "coarse_kernel", "fine": "fine_kernel"}[dom]
        if world == 1 and kname in tj["kernels"]:
            traffic = tj["kernels"][kname]["dram_bytes_read"] + tj["kernels"][kname]["dram_bytes_write"]
    except Exception:
        traffic = None
    dom_kernel = {"front": "flatten_subdivide_kernel + flatten_eseg_emit_kernel (+ classify, scans)", "binning": "path_count_kernel (+ backdrop, tiling)",
                  "coarse": "coarse_kernel (+ hit scan/scatter)", "fine": "fine_kernel"}[dom]

    # ---- e2e: public host API, host buffers, H2D + D2H inside the timed region
    acc = CUDAAccelerator(local_rank)
    acc.Init()
    acc.ctx.set_band(y0, y1)
    tgt = GPURenderTarget(w, h)
    for _ in range(2):
        acc.RenderEncoding(tgt, enc)
    if world > 1:
        dist.barrier()
    n_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        acc.RenderEncoding(tgt, enc)
    e2e_s = (time.perf_counter() - t0) / n_e2e
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = int(acc.ctx.stats()["scene_bytes"])
    d2h = band_h * w * 4
    acc.Close()

    line = {"metric": "Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "config3_4k_10k_paths_blend_layers_clips", "width": w, "height": h, "bands": world,
                       "band_assembly": assemble_kind, "rank_pipeline_ms_max": own_max_ms, "rank0_pipeline_ms": own_ms / args.steps,
                       "paths": int(st["n_draws"]), "l2": "flushed between timed iterations (256 MiB write)",
                       "frames_per_s": 1e3 / ms_per_step,
                       "stage_ms": {k: float(v / args.steps) for k, v in zip(("front", "binning", "coarse", "fine"), stage_ms)},
                       "counts": {k: int(st[k]) for k in ("n_lines", "n_path_tiles", "n_segments", "n_hits", "n_ptcl_words")}},
            "roofline": {"bound": "hbm", "kernel": dom_kernel, "stage": dom, "achieved": stages[dom]["achieved"], "peak": peak,
                         "unit": "GB/s", "frac": stages[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes": stages[dom]["algorithmic_bytes"], "kernel_ms": stages[dom]["ms"], "stages": stages,
                         "note": "every stage is issue/latency bound on this scene, not HBM bound; see profiles/"},
            "e2e": {"value": w * h / 1e6 / e2e_s, "unit": "Mpix/s", "ms_per_frame": e2e_s * 1e3,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks}

    if rank == 0 and world == 1 and args.cpu_baseline:
        from oracle import twin as T
        words_s = ctx.debug_read(_lib.BUF_SCENE, np.uint32)
        layout = ctx.debug_read(_lib.BUF_LAYOUT, _lib.LAYOUT)[0]
        t0 = time.perf_counter()
        _, tm = T.render_packed(words_s, layout, w, h, (0, 0, 0, 0), 1)
        cpu_s = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": w * h / 1e6 / cpu_s, "unit": "Mpix/s", "cores": 1, "kind": "port",
                                "sample": f"one whole {w}x{h} frame of the same scene, single thread ({cpu_s:.2f} s)",
                                "stage_s": {k: tm[k] for k in ("t_flatten", "t_coarse", "t_fine")}}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
