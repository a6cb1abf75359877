"""Horizontal band decomposition of the canvas across the GPUs of one box.

Tiles are independent after binning and backdrop only propagates along x inside a tile row, so bands of
whole tile rows have no cross-band dependency: every rank holds the whole scene, renders only its band
(ggcuda_set_band) and the bands are assembled with one all-gather. Equal bands let the collective be a single
`all_gather_into_tensor` straight into the frame (the canvas is padded to a multiple of world x 16 rows)."""
import torch
import torch.distributed as dist

TILE = 16


def band_rows(height_px, world, rank):
    """Tile rows [y0, y1) owned by `rank`; every rank gets the same count (the last ones may lie below the canvas)."""
    ht = (height_px + TILE - 1) // TILE
    per = (ht + world - 1) // world
    return rank * per, (rank + 1) * per


def padded_height(height_px, world):
    ht = (height_px + TILE - 1) // TILE
    per = (ht + world - 1) // world
    return per * world * TILE


def alloc_frame(width_px, height_px, world, device):
    """Frame buffer large enough for `world` equal bands; rows beyond height_px are padding."""
    return torch.zeros((padded_height(height_px, world), width_px, 4), dtype=torch.uint8, device=device)


def band_view(frame, height_px, world, rank):
    y0, y1 = band_rows(height_px, world, rank)
    return frame[y0 * TILE:y1 * TILE]


def assemble(frame, height_px, world, rank):
    """All-gather every rank's band into every rank's frame (in place: the band is a view of the frame)."""
    if world == 1:
        return frame
    band = band_view(frame, height_px, world, rank)
    dist.all_gather_into_tensor(frame.view(-1), band.reshape(-1))
    return frame


class SymmetricFrame:
    """The frame of every rank as ONE symmetric allocation (torch.distributed._symmetric_memory): each rank can address
    every other rank's copy (peer pointers over NVLink, and an NVSwitch multicast address where the fabric offers one),
    so fine rasterisation stores its band into all frames itself and the all-gather disappears; what is left of the
    collective is a barrier. `ptrs_for(rank)`: where `rank`'s band starts in every OTHER rank's frame."""

    def __init__(self, width_px, height_px, world, rank, device):
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank, self.height_px = world, rank, height_px
        self.frame = symm.empty((padded_height(height_px, world), width_px, 4), dtype=torch.uint8, device=device)
        self.frame.zero_()
        self.hdl = symm.rendezvous(self.frame, dist.group.WORLD)
        y0, _ = band_rows(height_px, world, rank)
        self.band_offset = y0 * TILE * width_px * 4
        mc = int(self.hdl.multicast_ptr or 0)   # 0 where the fabric / driver offers no multicast object
        self.multicast = mc != 0
        self.multicast_band = mc + self.band_offset if self.multicast else 0
        self.peer_bands = [int(p) + self.band_offset for r, p in enumerate(self.hdl.buffer_ptrs) if r != rank]

    def band(self):
        return band_view(self.frame, self.height_px, self.world, self.rank)

    def barrier(self):
        self.hdl.barrier()
