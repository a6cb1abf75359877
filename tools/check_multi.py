"""Multi-GPU band assembly check (run under torchrun, N >= 2): the frame every rank ends up with when fine stores its
band into all frames itself (multicast, then peer stores) must equal the frame assembled by the NCCL all-gather."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gg_b200 import _lib, bands, scenes  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
enc, w, h = scenes.config3(n=2000, bands=world)
ctx = _lib.Context(local)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
ctx.begin(w, h)
ctx.add_encoding(*enc.streams())
y0, y1 = bands.band_rows(h, world, rank)
ctx.set_band(y0, y1)
ctx.upload()
ref = bands.alloc_frame(w, h, world, "cuda")
ctx.render_device(bands.band_view(ref, h, world, rank).data_ptr(), w * 4, _lib.KEEP_SCENE)
bands.assemble(ref, h, world, rank)
torch.cuda.synchronize()
sym = bands.SymmetricFrame(w, h, world, rank, f"cuda:{local}")
results = {}
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
    results[mode] = bool((sym.frame == ref).all().item())
    dist.barrier()
nz = float((ref[..., 3] > 0).float().mean().item())
print(f"rank {rank}: multicast supported={sym.multicast} results={results} nonzero={nz:.3f}", flush=True)
ok = torch.tensor([int(all(results.values()))], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
ctx.close()
dist.destroy_process_group()
sys.exit(0 if ok.item() == 1 else 1)
