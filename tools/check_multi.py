"""Multi-GPU band assembly check (run under torchrun, N >= 2; tests/test_gpu_multi.py launches it). Ground truth: every rank
renders the WHOLE frame alone. Against it, bit for bit, on every rank:
  * bands (ingest-time culling active) assembled by the library's own NCCL all-gather (ggcuda_comm_init / ggcuda_all_gather_bands),
  * bands assembled by torch's all_gather_into_tensor,
  * bands stored by the fine kernel itself into every rank's frame: NVSwitch multicast (multimem.st), then peer pointers,
    each followed by the symmetric-memory barrier only,
  * bands rendered privately and broadcast afterwards on a side stream (ggcuda_broadcast_band), multicast and peer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gg_b200 import _lib, bands, scenes  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_paths = int(os.environ.get("GG_CHECK_PATHS", "2000"))
enc, w, h = scenes.config3(n=n_paths, bands=world)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)

# ground truth: the whole frame on this device alone
solo = _lib.Context(local)
solo.set_stream(stream.cuda_stream)
solo.begin(w, h)
solo.add_encoding(*enc.streams())
truth = bands.alloc_frame(w, h, world, "cuda")
solo.render_device(truth.data_ptr(), w * 4, 0)
torch.cuda.synchronize()
n_all = solo.stats()["n_draws"]
solo.close()

ctx = _lib.Context(local)
ctx.set_stream(stream.cuda_stream)
y0, y1 = bands.band_rows(h, world, rank)
ctx.begin(w, h)
ctx.set_band(y0, y1)                 # before the scene: paths that cannot reach the band are dropped at ingest
ctx.add_encoding(*enc.streams())
ctx.upload()
results = {}

# 1. the library's own all-gather
uid = [_lib.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(world, rank, uid[0])
frame = bands.alloc_frame(w, h, world, "cuda")
band = bands.band_view(frame, h, world, rank)
ctx.render_device(band.data_ptr(), w * 4, _lib.KEEP_SCENE)
ctx.all_gather_bands(frame.data_ptr(), band.numel())
ctx.sync()
torch.cuda.synchronize()
results["lib_nccl"] = bool((frame == truth).all().item())
culled = n_all - ctx.stats()["n_draws"]

# 2. torch's all-gather
frame.zero_()
ctx.render_device(band.data_ptr(), w * 4, _lib.KEEP_SCENE)
bands.assemble(frame, h, world, rank)
torch.cuda.synchronize()
results["torch_nccl"] = bool((frame == truth).all().item())

# 3. fused stores
try:
    sym = bands.SymmetricFrame(w, h, world, rank, f"cuda:{local}")
except Exception as e:   # noqa: BLE001
    sym = None
    print(f"rank {rank}: symmetric memory unavailable: {type(e).__name__}: {e}", flush=True)
if sym is not None:
    for mode in (["multicast"] if sym.multicast else []) + ["peer"]:
        sym.frame.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        if mode == "multicast":
            ctx.render_device_multi(sym.band().data_ptr(), [sym.multicast_band], w * 4, _lib.KEEP_SCENE, multicast=True)
        else:
            ctx.render_device_multi(sym.band().data_ptr(), sym.peer_bands, w * 4, _lib.KEEP_SCENE)
        sym.barrier()
        torch.cuda.synchronize()
        results[mode] = bool((sym.frame == truth).all().item())
        dist.barrier()
    # 4. deferred broadcast: three frames rendered alternately into two private bands, each copied into every rank's frame
    #    on a side stream (ggcuda_broadcast_band) while the next one is rasterised; barrier on the side stream
    side = torch.cuda.Stream()
    scratch = [torch.empty_like(sym.band()), torch.empty_like(sym.band())]
    nbytes = scratch[0].numel()
    for mode in (["bcast_multicast"] if sym.multicast else []) + ["bcast_peer"]:
        sym.frame.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        for k in range(3):
            sb = scratch[k & 1]
            sb.zero_()
            ctx.render_device(sb.data_ptr(), w * 4, _lib.KEEP_SCENE | _lib.NO_WAIT)   # queued behind one another, no host wait
            if mode == "bcast_multicast":
                ctx.broadcast_band(sb.data_ptr(), [sym.multicast_band], nbytes, side.cuda_stream, multicast=True)
            else:
                ctx.broadcast_band(sb.data_ptr(), sym.peer_bands + [sym.band().data_ptr()], nbytes, side.cuda_stream)
            with torch.cuda.stream(side):
                sym.barrier()
        torch.cuda.synchronize()
        results[mode] = bool((sym.frame == truth).all().item())
        dist.barrier()
nz = float((truth[..., 3] > 0).float().mean().item())
print(f"rank {rank}: results={results} culled_draws={culled} of {n_all} nonzero={nz:.3f}", flush=True)
ok = torch.tensor([int(all(results.values()) and nz > 0.1 and (culled > 0 or world == 1))], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
ctx.comm_destroy()
ctx.close()
dist.destroy_process_group()
if rank == 0:
    print("CHECK_MULTI " + ("OK" if ok.item() == 1 else "FAILED"), flush=True)
sys.exit(0 if ok.item() == 1 else 1)
