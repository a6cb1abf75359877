#!/usr/bin/env python
"""bench.py -- Mpixel/s of gg's scene-rasterisation hot path on B200 (ggcuda) and on the host CPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" is one pass of the hot path (packed scene already in HBM -> flatten incl. stroke expansion -> bin ->
coarse/PTCL -> fine -> RGBA8 band in HBM, then the assembly of the bands in every rank's frame when N > 1).

Workloads (BASELINE.json configs): config1 (512^2, 1 000 fills), config2 (1920x1080 SVG icons), config3 (default: 3840x2160,
10 000 filled + stroked Bezier paths, 29 blend modes, layers, clips -- the configuration the metric is quoted on),
config3_nowipe (the same with the 23 blend modes that never blank a tile), config4 (4K, 50 000 glyph outlines), config5
(16384^2, 1 M paths). N GPUs: config3 is WEAK scaling -- N copies of the 4K frame stacked vertically, one band of 135 tile
rows per GPU; every other workload is STRONG scaling -- the one canvas cut into N bands. A rank ingests only the paths
that can reach its band. Band assembly (GG_BANDS): `p2p` -- the fine kernel stores its band into every rank's frame itself
(symmetric memory: one multimem.st per 16 bytes through the NVSwitch, `p2p_nomc`: peer stores) followed by a barrier;
`p2p_async` -- the band is rendered privately and broadcast on a side stream while the next frame is rasterised
(ggcuda_broadcast_band); `nccl` / `torch_nccl` -- one all-gather issued by the library / by torch. Default `auto`: p2p, or
p2p_async when the receivers' ingress ((N - 1) bands per frame) would take longer than fine itself; NCCL if symmetric memory
cannot be set up. A pass into device memory is replayed as one CUDA graph (GGCUDA_NO_GRAPH=1: through the stream). `value` is
device time (CUDA events, max over ranks; in p2p_async ONE event pair around all K steps, L2 flushes included); the per-stage
times come from a second set of passes with events between the stages. `e2e` is the same frame through the public host API
with host buffers (scene ingest + H2D + pipeline + D2H inside the timed region). At N > 1 every rank's assembled frame is checked against the bands the ranks
rendered (`frame_ok`). `--impl reference` times the CPU restatement of the reference's pipeline on the host cores.
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

WORKLOADS = {
    "config1": "config1_512_1k_fills",
    "config2": "config2_1080p_svg_icons",
    "config3": "config3_4k_10k_paths_blend_layers_clips",
    "config3_nowipe": "config3_4k_10k_paths_23_nonwiping_blend_modes",
    "config4": "config4_4k_50k_glyph_outlines",
    "config5": "config5_16k_1m_paths",
}


def build_workload(name, bands=1):
    """(encoding, width, height, background premul RGBA8, scaling)."""
    from gg_b200 import scenes
    if name == "config1":
        return (*scenes.config1(), (0, 0, 0, 0), "strong")
    if name == "config2":
        return (*scenes.config2(), (255, 255, 255, 255), "strong")
    if name == "config3":
        return (*scenes.config3(bands=bands), (0, 0, 0, 0), "weak")
    if name == "config3_nowipe":
        return (*scenes.config3(bands=bands, nowipe=True), (0, 0, 0, 0), "weak")
    if name == "config4":
        glyphs = json.load(open(os.path.join(ROOT, "tests", "golden", "fixtures", "goregular_ascii.json")))
        return (*scenes.config4(glyphs), (255, 255, 255, 255), "strong")
    if name == "config5":
        return (*scenes.config5(), (0, 0, 0, 0), "strong")
    raise SystemExit(f"unknown workload {name!r}; choose from {sorted(WORKLOADS)}")


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


def fine_bytes_fast(ptcl_off, ptcl, w, h):
    """B_fine = 4 W H + 4 sum(ptcl words) + 20 sum over tile-fills of seg_count, walking all tiles in
    lock-step from the word offsets given (list start, or the restart point)."""
    pos = ptcl_off.astype(np.int64) + 1
    active = np.ones(len(pos), dtype=bool)
    words = len(pos)
    segs = 0
    size = np.zeros(16, dtype=np.int64)
    size[[0, 1, 3, 5, 6, 10, 11]] = [1, 4, 1, 2, 2, 1, 3]
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
    """The reference's CPU implementation of the path (oracle port of internal/gpu/tilecompute; the Go original cannot be
    built here: no Go toolchain), all host threads. Bounded sample: ONE band of this arm's workload per step -- for config3
    that is one whole 3840x2160 frame whatever N is (the per-GPU share of the weak-scaling canvas); the line's config
    says what was rendered, not what the N-GPU arm renders."""
    if rank != 0:
        return
    from oracle import twin as T
    from gg_b200 import _lib
    name = args.workload
    if name == "config5":   # 1/64 of the 16K canvas: the top-left 2048 x 2048 of the same encoding (a smaller canvas only clips)
        enc, w, h, bg, scaling = build_workload(name)
        w = h = 2048
        sample = "top-left 2048x2048 of the 16384^2 canvas (the oracle's coarse is O(tiles x draws): the whole canvas does not finish)"
    else:
        enc, w, h, bg, scaling = build_workload(name, bands=1)
        sample = f"one whole {w}x{h} frame per step"
    hc = _lib.Context(-1)
    hc.begin(w, h)
    hc.add_encoding(*enc.streams())
    words, layout = hc.pack_host()
    hc.close()
    threads = os.cpu_count() or 1
    for _ in range(min(args.warmup, 2)):
        T.render_packed(words, layout, w, h, bg, threads)
    t0 = time.perf_counter()
    tm = {}
    for _ in range(args.steps):
        _, tm = T.render_packed(words, layout, w, h, bg, threads)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    val = w * h / 1e6 / dt
    line = {"impl": "reference", "metric": "Mpix/s", "value": val, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[name], "width": w, "height": h, "bands": 1,
                       "note": "CPU arm: one band on the host cores, whatever --gpus says; own C port of internal/gpu/tilecompute, not gg's Go code"},
            "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} (flatten+coarse on 1 thread, fine on {threads} threads); "
                                       "port of internal/gpu/tilecompute (the Go original cannot be built here)",
                             "stage_s": {k: tm.get(k) for k in ("t_flatten", "t_coarse", "t_fine")}},
            "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The one JSON line, on the process's real stdout."""
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def checksum_bands(frame, n_bands):
    """One position-weighted 64-bit checksum per band of an assembled frame."""
    import torch
    v = frame.reshape(n_bands, -1).view(torch.int32).to(torch.int64)
    wgt = (torch.arange(v.shape[1], device=v.device, dtype=torch.int64) % 65521) + 1
    return (v * wgt).sum(dim=1)


def time_pipeline(ctx, torch, stream, steps, warmup, step_fn, flush_buf, dist=None, stages=True):
    """(ms per step summed, per-stage ms summed, own-pipeline ms summed) over `steps` timed iterations, L2 flushed between."""
    for _ in range(warmup):
        step_fn()
    torch.cuda.synchronize()
    total_ms, stage_ms, own_ms = 0.0, np.zeros(4), 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(steps):
        flush_buf.fill_(1)          # L2 flush between timed iterations (not timed)
        e0.record(stream)
        e_mid = step_fn()
        e1.record(stream)
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        own_ms += e0.elapsed_time(e_mid)
        if stages:
            s = ctx.stats()
            stage_ms += [s["ms_front"], s["ms_binning"], s["ms_coarse"], s["ms_fine"]]
    return total_ms, stage_ms, own_ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-baseline", type=int, default=1, help="time the oracle on the host cores at N=1 (0 to skip)")
    ap.add_argument("--variants", type=int, default=1, help="at N=1 also time config3_nowipe beside config3 (0 to skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # Watchdog: a wedged collective or driver call must not hold a GPU box until the caller's limit; dump every
    # thread's stack and exit instead.
    faulthandler.dump_traceback_later(int(os.environ.get("GG_BENCH_WATCHDOG_S", "900")), exit=True)
    # stdout carries exactly ONE line, the JSON: whatever a library prints there meanwhile (NCCL's version banner does) goes
    # to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from gg_b200 import _lib, bands
    from gg_b200.accelerator import CUDAAccelerator, GPURenderTarget

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    args.warmup = max(3, args.warmup)

    enc, w, h, bg, scaling = build_workload(args.workload, bands=world)
    streams = enc.streams()
    y0, y1 = bands.band_rows(h, world, rank)

    ctx = _lib.Context(local_rank)
    stream = torch.cuda.Stream()          # kernels, events and the all-gather all go through this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_timing(True)
    ctx.begin(w, h)
    ctx.set_background(bg)
    ctx.set_band(y0, y1)                  # before the scene: paths that cannot reach the band are dropped at ingest
    ctx.add_encoding(*streams)
    ctx.upload()
    # Full canvas on every rank. p2p: a symmetric allocation -- fine stores its band into every rank's frame itself (NVSwitch
    # multicast where offered, else peer stores) and a barrier replaces the all-gather. nccl: one all-gather issued by the
    # library on its own communicator (ggcuda_all_gather_bands); torch_nccl: torch's. NCCL is also the fallback if
    # symmetric memory cannot be set up on this box.
    sym, assemble_kind = None, "single"
    band_mode = os.environ.get("GG_BANDS", "auto")
    if band_mode == "auto":
        # Fused stores (fine writes its band into every frame) unless the receivers' NVLink ingress -- (N - 1) bands per
        # device and frame -- takes longer than fine itself: then fine would be stretched to the ingress time, and the band
        # is broadcast afterwards on a side stream, behind the next frame's front stages. Measured on B200s, round 2,
        # config3: N = 8 fused 1.47 ms / step, deferred 1.41 (L2 flush inside), NCCL all-gather 2.14; N = 4 fused 1.31,
        # deferred 1.32; config5 at N = 8 (fine 3.6 ms, ingress 1.3 ms): fused 6.0, deferred 6.5.
        band_mode = "p2p"
        if world > 1:
            ctx.set_timing(True)
            probe = torch.zeros((max(1, min(y1 * 16, h) - y0 * 16), w, 4), dtype=torch.uint8, device="cuda")
            for _ in range(3):
                ctx.render_device(probe.data_ptr(), w * 4, _lib.KEEP_SCENE)
            fine_ms = torch.tensor([ctx.stats()["ms_fine"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(fine_ms, op=dist.ReduceOp.MAX)
            ingress_ms = (world - 1) * probe.numel() / 700e9 * 1e3      # ~700 GB/s of the 900 GB/s a direction offers
            if ingress_ms > float(fine_ms.item()):
                band_mode = "p2p_async"
            del probe
            ctx.set_timing(False)
    deferred = band_mode in ("p2p_async", "p2p_async_nomc")   # bands broadcast on a side stream, behind the next frame's front stages
    if world > 1 and band_mode in ("p2p", "p2p_nomc", "p2p_async", "p2p_async_nomc"):
        try:
            sym = bands.SymmetricFrame(w, h, world, rank, f"cuda:{local_rank}")
            if band_mode in ("p2p_nomc", "p2p_async_nomc"):
                sym.multicast = False
            via = "multimem.st over NVSwitch multicast" if sym.multicast else "peer memory over NVLink"
            assemble_kind = (f"band broadcast into all frames ({via}) on a side stream while the next frame is rasterised (ggcuda_broadcast_band) + barrier"
                             if deferred else f"fine stores into all frames ({via}) + barrier")
        except Exception as e:   # noqa: BLE001
            if rank == 0:
                print(f"symmetric memory unavailable ({type(e).__name__}: {e}); using the NCCL all-gather", file=sys.stderr)
            sym = None
        ok = torch.tensor([1 if sym is not None else 0], device="cuda")   # every rank must have it, or nobody uses it
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            sym, band_mode = None, "nccl"
    lib_nccl = False
    if sym is not None:
        frame, band = sym.frame, sym.band()
    else:
        frame = bands.alloc_frame(w, h, world, "cuda")
        band = bands.band_view(frame, h, world, rank)     # fine writes its band straight into the gather buffer
        if world > 1 and band_mode != "torch_nccl":
            uid = [_lib.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.comm_init(world, rank, uid[0])
            lib_nccl = True
            assemble_kind = "NCCL all-gather issued by libggcuda (ggcuda_all_gather_bands)"
        elif world > 1:
            assemble_kind = "NCCL all_gather_into_tensor (torch)"
    stride = w * 4
    band_bytes = band.numel()
    flush_buf = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # 1.5 x the 126 MB L2
    e_mid = torch.cuda.Event(enable_timing=True)

    side = torch.cuda.Stream() if (sym is not None and deferred) else None
    scratch = [torch.empty_like(band), torch.empty_like(band)] if side is not None else None
    step_no = [0]

    def make_step(c, band_t):
        def step():
            if side is not None:
                # render into one of two private bands; a small kernel on the side stream copies it into every rank's frame
                # (this rank's too) and the barrier follows it there, while this stream goes on with the next frame
                sb = scratch[step_no[0] & 1]
                step_no[0] += 1
                c.render_device(sb.data_ptr(), stride, _lib.KEEP_SCENE | _lib.NO_WAIT)   # the host queues frames ahead of the device
                e_mid.record(stream)
                if sym.multicast:
                    c.broadcast_band(sb.data_ptr(), [sym.multicast_band], band_bytes, side.cuda_stream, multicast=True)
                else:
                    c.broadcast_band(sb.data_ptr(), sym.peer_bands + [band_t.data_ptr()], band_bytes, side.cuda_stream)
                with torch.cuda.stream(side):
                    sym.barrier()
            elif sym is not None:
                if sym.multicast:
                    c.render_device_multi(band_t.data_ptr(), [sym.multicast_band], stride, _lib.KEEP_SCENE, multicast=True)
                else:
                    c.render_device_multi(band_t.data_ptr(), sym.peer_bands, stride, _lib.KEEP_SCENE)
                e_mid.record(stream)
                sym.barrier()
            else:
                c.render_device(band_t.data_ptr(), stride, _lib.KEEP_SCENE)
                e_mid.record(stream)
                if lib_nccl:
                    c.all_gather_bands(frame.data_ptr(), band_bytes)
                elif world > 1:
                    bands.assemble(frame, h, world, rank)
            return e_mid
        return step

    step = make_step(ctx, band)
    # The timed region runs without per-stage events: a pass into a device target is then one CUDA-graph replay
    # (GGCUDA_NO_GRAPH=1 turns that off). The per-stage times come from a second, untimed set of passes below.
    ctx.set_timing(False)
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
    if side is not None:
        # pipelined across frames: one event pair around all K steps (the L2 flushes between them are inside it and counted),
        # closed only after the side stream has delivered the last band everywhere
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            flush_buf.fill_(1)
            step()
        stream.wait_stream(side)
        e1.record(stream)
        torch.cuda.synchronize()
        total_ms, own_ms = e0.elapsed_time(e1), 0.0
    else:
        total_ms, _, own_ms = time_pipeline(ctx, torch, stream, args.steps, 0, step, flush_buf, stages=False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ctx.set_timing(True)     # stage breakdown: the same passes again with CUDA events between the stages (stream launches)
    staged_total_ms, stage_ms, _ = time_pipeline(ctx, torch, stream, args.steps, 2, step, flush_buf)
    ctx.set_timing(False)
    t = torch.tensor([total_ms, own_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    own_max_ms = float(t[1].item()) / args.steps   # slowest rank's own pipeline, before it waits for the others
    if side is not None:
        own_max_ms = None                          # frames overlap in this mode: there is no per-frame boundary to time
    ms_per_step = float(t[0].item()) / args.steps
    value = w * h / 1e6 / (ms_per_step / 1e3)

    # ---- N > 1: is the frame every rank holds the frame the ranks rendered? Each rank renders its band once more into a
    #      private buffer; position-weighted checksums of those bands (all-gathered as 8-byte values) against the same
    #      checksums over the assembled frame of every rank.
    frame_ok = None
    if world > 1:
        private = torch.zeros_like(band)
        ctx.render_device(private.data_ptr(), stride, _lib.KEEP_SCENE)
        torch.cuda.synchronize()
        mine = checksum_bands(private, 1)
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        want = torch.cat(allc)
        got = checksum_bands(frame, world)
        okt = torch.tensor([int(bool((want == got).all().item()) and bool((private != 0).any().item()))], device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        frame_ok = bool(okt.item())

    # ---- roofline: algorithmic bytes / CUDA-event time, per stage; the headline object is
    #      the stage that takes longest. Fine's bytes count only what lies after each tile's restart point (the
    #      part of the PTCL it has to execute), so skipping dead commands does not inflate its GB/s.
    peak, peak_src = _peaks()
    band_h = max(0, min(y1 * 16, h) - y0 * 16)

    def roofline_of(c, stage_ms_sum, steps):
        poff = c.debug_read(_lib.BUF_PTCL_OFF, np.uint32)
        ptcl = c.debug_read(_lib.BUF_PTCL, np.uint32)
        rst = c.debug_read(_lib.BUF_RESTART, np.uint32).reshape(-1, 2)[:, 0]
        lay = c.debug_read(_lib.BUF_LAYOUT, _lib.LAYOUT)[0]
        s_ = c.stats()
        _, words_all, segs_all = fine_bytes_fast(poff, ptcl, w, band_h)
        b_fine, words, segs = fine_bytes_fast(poff + np.maximum(rst, 1) - 1, ptcl, w, band_h)
        n_pd = int(lay["draw_tag_base"] - lay["path_data_base"])
        n_tr = int(lay["style_base"] - lay["transform_base"]) // 6
        cnt = {k: int(s_[k]) for k in ("n_lines", "n_path_tiles", "n_segments", "n_hits", "n_ptcl_words", "n_draws", "n_tag_bytes")}
        stage_bytes = {
            "front": cnt["n_tag_bytes"] + 4 * n_pd + 24 * n_tr + 20 * cnt["n_lines"] + 6 * cnt["n_tag_bytes"] + 20 * cnt["n_draws"],
            "binning": 20 * cnt["n_lines"] + 8 * cnt["n_segments"] + 8 * cnt["n_path_tiles"] + 32 * cnt["n_path_tiles"] + 48 * cnt["n_segments"],
            "coarse": 16 * cnt["n_draws"] + 8 * cnt["n_path_tiles"] + 4 * words_all,
            "fine": b_fine,
        }
        stages = {}
        for k, ms in zip(("front", "binning", "coarse", "fine"), stage_ms_sum / steps):
            gbs = stage_bytes[k] / (ms / 1e3) / 1e9 if ms > 0 else 0.0
            stages[k] = {"ms": float(ms), "algorithmic_bytes": int(stage_bytes[k]), "achieved": gbs, "frac": gbs / peak}
        fine_info = {"ptcl_words": int(words_all), "ptcl_words_executed": int(words), "restart_skipped_frac": 1.0 - words / max(1, words_all),
                     "segments_in_fills": int(segs_all), "segments_evaluated": int(segs)}
        return stages, cnt, fine_info

    stages, counts, fine_info = roofline_of(ctx, stage_ms, args.steps)
    dom = max(stages, key=lambda k: stages[k]["ms"])
    # DRAM bytes of the dominant stage's main kernel from the committed `ncu --set full` capture of this workload
    # (profiles/*_traffic.json; null if that kernel was not captured)
    traffic = None
    try:
        tj = json.load(open(sorted(p for p in (os.path.join(ROOT, "profiles", f) for f in os.listdir(os.path.join(ROOT, "profiles")))
                                   if p.endswith("_traffic.json"))[-1]))
        kname = {"front": "flatten_subdivide_kernel", "binning": "path_count_kernel", "coarse": "coarse_kernel", "fine": "fine_kernel"}[dom]
        if world == 1 and args.workload == "config3" and kname in tj["kernels"]:
            traffic = tj["kernels"][kname]["dram_bytes_read"] + tj["kernels"][kname]["dram_bytes_write"]
    except Exception:
        traffic = None
    dom_kernel = {"front": "flatten_subdivide_kernel + flatten_eseg_emit_kernel (+ classify, scans)", "binning": "path_count_kernel (+ backdrop, tiling)",
                  "coarse": "coarse_kernel (+ hit scan/scatter)", "fine": "fine_kernel"}[dom]

    # ---- the same scene without the blend modes that blank tiles (N = 1, config3 only): what fine costs when restart
    #      points only come from opaque fills
    variant = None
    if world == 1 and args.workload == "config3" and args.variants:
        enc_v, wv, hv, bgv, _ = build_workload("config3_nowipe")
        cv = _lib.Context(local_rank)
        cv.set_stream(stream.cuda_stream)
        cv.set_timing(True)
        cv.begin(wv, hv)
        cv.add_encoding(*enc_v.streams())
        cv.upload()
        nsteps = max(3, min(args.steps, 10))
        tv, sv, _ = time_pipeline(cv, torch, stream, nsteps, 3, make_step(cv, band), flush_buf)
        stv, _, fiv = roofline_of(cv, sv, nsteps)
        variant = {"workload": WORKLOADS["config3_nowipe"], "ms_per_step": tv / nsteps, "Mpix_per_s": wv * hv / 1e6 / (tv / nsteps / 1e3),
                   "stage_ms": {k: v["ms"] for k, v in stv.items()}, "fine": {**fiv, "achieved_GBps": stv["fine"]["achieved"], "frac": stv["fine"]["frac"]}}
        cv.close()

    # ---- e2e: public host API (the scene.EncodingAccelerator entry), host buffers, scene ingest + H2D + pipeline + D2H inside
    #      the timed region. The target's pixels are page-locked once (what the Go binding does for a Context's pixmap).
    #      `resident`: the same call when the encoding's key is still resident on the device (fine + read-back only).
    acc = CUDAAccelerator(local_rank)
    acc.Init()
    acc.ctx.set_band(y0, y1)
    acc.ctx.set_background(bg)
    tgt = GPURenderTarget(w, h)
    acc.PinTarget(tgt)
    for _ in range(2):
        acc.RenderEncoding(tgt, enc, resident=False)
    if world > 1:
        dist.barrier()
    n_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        acc.RenderEncoding(tgt, enc, resident=False)
    e2e_s = (time.perf_counter() - t0) / n_e2e
    h2d = int(acc.ctx.stats()["scene_bytes"])
    acc.RenderEncoding(tgt, enc)          # leaves the scene resident under its key
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        acc.RenderEncoding(tgt, enc)
    res_s = (time.perf_counter() - t0) / n_e2e
    assert acc.resident_hits >= n_e2e
    # the same cold frames with TWO in flight (two contexts, two host threads): ingest of frame k + 1 beside the device's frame k
    acc.Close()
    from gg_b200.accelerator import PipelinedRenderer
    pr = PipelinedRenderer(local_rank, depth=2, band=(y0, y1), background=bg)
    tgts = [GPURenderTarget(w, h), GPURenderTarget(w, h)]
    for a, t_ in zip(pr.accs, tgts):
        a.PinTarget(t_)
    futs = [pr.submit(tgts[k & 1], enc, resident=False) for k in range(4)]
    [f.result() for f in futs]
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    futs = [pr.submit(tgts[k & 1], enc, resident=False) for k in range(2 * n_e2e)]
    [f.result() for f in futs]
    pipe_s = (time.perf_counter() - t0) / (2 * n_e2e)
    pipe_ok = bool((tgts[0].Data == tgt.Data).all() and (tgts[1].Data == tgt.Data).all())
    pr.Close()
    te = torch.tensor([e2e_s, res_s, pipe_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, res_s, pipe_s = float(te[0].item()), float(te[1].item()), float(te[2].item())
    d2h = band_h * w * 4

    line = {"metric": "Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "width": w, "height": h, "bands": world,
                       "band_assembly": assemble_kind, "rank_pipeline_ms_max": own_max_ms, "rank0_pipeline_ms": (own_ms / args.steps if side is None else None),
                       "paths": int(st["n_draws"]), "l2": "flushed between timed iterations (192 MiB write)" + (", inside the timed region in this mode" if side is not None else ""),
                       "frames_per_s": 1e3 / ms_per_step,
                       "stage_ms": {k: float(v / args.steps) for k, v in zip(("front", "binning", "coarse", "fine"), stage_ms)},
                       "counts": {k: counts[k] for k in ("n_lines", "n_path_tiles", "n_segments", "n_hits", "n_ptcl_words")},
                       "stage_ms_note": "from a second set of passes with events between the stages (%.3f ms/step there); the timed passes replay one CUDA graph" % (staged_total_ms / args.steps),
                       "cuda_graph": os.environ.get("GGCUDA_NO_GRAPH", "0") in ("", "0"),
                       "fine": fine_info, "launches_per_step": int(launches_per_step)},
            "roofline": {"bound": "hbm", "kernel": dom_kernel, "stage": dom, "achieved": stages[dom]["achieved"], "peak": peak,
                         "unit": "GB/s", "frac": stages[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes": stages[dom]["algorithmic_bytes"], "kernel_ms": stages[dom]["ms"], "stages": stages,
                         "note": "stages are issue/latency bound on this scene, not HBM bound; see profiles/"},
            "e2e": {"value": w * h / 1e6 / e2e_s, "unit": "Mpix/s", "ms_per_frame": e2e_s * 1e3,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h),
                    "resident": {"ms_per_frame": res_s * 1e3, "Mpix_per_s": w * h / 1e6 / res_s, "h2d_bytes_per_step": 0,
                                 "note": "same call, encoding key still resident on the device: fine + read-back only"},
                    "pipelined": {"ms_per_frame": pipe_s * 1e3, "Mpix_per_s": w * h / 1e6 / pipe_s, "frames_in_flight": 2, "frames_ok": pipe_ok,
                                  "note": "the cold call (ingest + H2D + pipeline + D2H every frame) from two host threads on two contexts: "
                                          "frame k + 1 is ingested while frame k is on the device"}},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks}
    if frame_ok is not None:
        line["frame_ok"] = frame_ok
    if variant is not None:
        line["config"]["nowipe_variant"] = variant

    if rank == 0 and world == 1 and args.cpu_baseline:
        from oracle import twin as T
        if args.workload == "config5":
            cw = ch = 2048
            hc = _lib.Context(-1)
            hc.begin(cw, ch)
            hc.add_encoding(*streams)
            words_s, layout = hc.pack_host()
            hc.close()
            sample = "top-left 2048x2048 of the 16384^2 canvas, single thread"
        else:
            cw, ch = w, h
            words_s = ctx.debug_read(_lib.BUF_SCENE, np.uint32)
            layout = ctx.debug_read(_lib.BUF_LAYOUT, _lib.LAYOUT)[0]
            sample = f"one whole {w}x{h} frame of the same scene, single thread"
        t0 = time.perf_counter()
        _, tm = T.render_packed(words_s, layout, cw, ch, bg, 1)
        cpu_s = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": cw * ch / 1e6 / cpu_s, "unit": "Mpix/s", "cores": 1, "kind": "port",
                                "sample": f"{sample} ({cpu_s:.2f} s); own C port of internal/gpu/tilecompute, not gg's Go code",
                                "stage_s": {k: tm[k] for k in ("t_flatten", "t_coarse", "t_fine")}}
    if rank == 0:
        emit(line)
    if lib_nccl:
        ctx.comm_destroy()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if frame_ok is False:
        sys.exit(3)


if __name__ == "__main__":
    main()
