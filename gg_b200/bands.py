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
