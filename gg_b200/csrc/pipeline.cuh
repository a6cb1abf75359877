// gg_b200/csrc/pipeline.cuh -- device buffer table shared by the host API and the stage launchers.
#pragma once
#include "common.cuh"

// All device pointers of one context. Capacities live in GGConfig (passed by value to kernels).
struct GGBuffers {
    // inputs
    uint32_t* scene;            // packed scene (tags | path data | draw tags | draw data | transforms | styles | clip aux)
    // scans
    GGPathMonoid* tag_monoids;  // [n_tag_words]  exclusive
    GGDrawMonoid* draw_monoids; // [n_draws]      exclusive (+ clip_leaf fix-up)
    uint32_t* info;             // [n_draws]      draw_leaf info (packed colours)
    GGClipInp* clip_inps;       // [n_clips]
    GGDrawRec* draw_recs;       // [n_draws]
    // flatten
    uint32_t* line_count;       // [n_tag_bytes]
    uint32_t* line_off;         // [n_tag_bytes]
    uint32_t* curve_list;       // [n_tag_bytes] tag-byte indices of quad / cubic tags (dense work list)
    struct GGESeg* esegs;       // [esegs_cap] Euler-segment records (flatten.cuh)
    GGLine* lines;              // [lines_cap]
    uint32_t* path_bbox_ord;    // [4*n_paths] order-mapped float min/max
    // binning
    GGPath* paths;              // [n_paths]
    uint32_t* path_row_off;     // [n_paths]
    GGTile* tiles;              // [tiles_cap]
    uint32_t* seg_start;        // [tiles_cap]
    uint8_t* imp_mask;          // [tiles_cap] per path tile: which implicit ancestors it brought into its tile's hit list
    uint32_t* imp_seen;         // [n_implicit * imp_words] one bit per (implicit layer, band tile)
    GGSegCount* seg_counts;     // [seg_counts_cap]
    GGSegment* segments;        // [segments_cap]
    // coarse
    unsigned long long* tile_hits;  // [band_tiles] (hits | words<<32) histogram
    uint32_t* hit_off;          // [band_tiles]
    uint32_t* hit_cnt;          // [band_tiles]
    uint32_t* hit_cursor;       // [band_tiles]
    GGHit* hits;                // [hits_cap] per-tile hit lists (draw index + path-tile record)
    uint32_t* ptcl_off;         // [band_tiles] (multiples of 4 words: 16-byte aligned lists for bulk copies)
    uint32_t* ptcl_len;         // [band_tiles] words actually written (incl. word 0 and CmdEnd)
    uint32_t* ptcl;             // [ptcl_cap]
    uint32_t* restart_pt;       // [2 * band_tiles] {PTCL word offset fine may start from (0 = list start), RGBA8 all pixels hold there}
    uint32_t* spill_off;        // [band_tiles] blend spill offsets (tile-levels), 0xffffffff = none; then [band_tiles] heavy tile list (bump->heavy)
    float4* spill;              // [spill_cap * 256]
    // misc
    GGBump* bump;
    void* scan_partials;        // GG_SCAN_BLOCKS * 32 bytes
};

// stage launchers (pipeline.cu)
// each returns the number of kernels it launched
uint32_t gg_launch_front(const GGConfig& cfg, const GGBuffers& b, cudaStream_t s);   // scans, flatten, path setup
uint32_t gg_launch_binning(const GGConfig& cfg, const GGBuffers& b, cudaStream_t s); // path_count, backdrop, seg alloc, path_tiling
uint32_t gg_launch_coarse(const GGConfig& cfg, const GGBuffers& b, cudaStream_t s);  // hit lists + PTCL
// fine (fine.cu). dst: RGBA8 premultiplied, row stride in bytes; covers tile rows [band_y0, band_y1).
// Further destinations of fine's band besides `dst`: the same band inside the frames of the other devices of a
// multi-GPU group (peer pointers, same row stride), or ONE multicast address that reaches all of them.
#define GG_MAX_MIRRORS 15
struct GGFineMirrors { uint8_t* p[GG_MAX_MIRRORS]; uint32_t n; uint32_t multicast; };
// What one fine launch covers: tile rows [row0, row1) relative to the band, tile-PAIR columns [px0, px1) (a warp owns two
// horizontally adjacent tiles). A dirty rectangle (resident scenes) narrows both.
struct GGFineRange { uint32_t row0, row1, px0, px1; };
void gg_launch_band_bcast(const void* band, const GGFineMirrors& mir, size_t bytes, uint32_t sm_count, cudaStream_t s);
void gg_launch_fine(const GGConfig& cfg, const GGBuffers& b, uint8_t* dst, size_t stride, cudaStream_t s, const GGFineRange& rg, uint32_t part,
                    const GGFineMirrors& mir);
