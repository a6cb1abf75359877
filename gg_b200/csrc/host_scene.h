// gg_b200/csrc/host_scene.h -- host-side scene packing for libggcuda.
//
// Turns gg's two input forms into the packed, Vello-style scene the device consumes:
//   * per-draw paths (gg.Path verbs + f64 device-space coords), the GPUAccelerator.FillPath /
//     StrokePath boundary (accelerator.go:118-128, path_convert.go:29-112), and
//   * whole scene.Encoding streams (scene/encoding.go:407-444) with the tag semantics of
//     scene.Renderer.executeEncodingOnTile (scene/renderer.go:619-813).
// Packed layout == gg's PackedScene (tilecompute/scene_encode.go:52-66, 280-356):
//   path tags (4 per u32, padded to 256 words) | path data | draw tags | draw data |
//   transforms | styles, followed by our clip-aux words and five constant words.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>

// Output of the host stroke expander: closed outline loops as MoveTo/LineTo tags + coordinates (device space).
struct StrokeSink {
    std::vector<uint8_t> tags;
    std::vector<float> data;
    uint32_t n_seg = 0;
    float bb[4] = {3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f};
};

struct StrokeStyleHost;
struct HostScene {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> tags;
    std::vector<float> path_data;
    std::vector<uint32_t> draw_tags, draw_data, styles;   // styles: 3 words per path {flags, width bits, miter-limit bits}
    bool host_strokes = false;           // diagnostic: expand strokes with the host polyline stroker instead of on the device
    std::vector<float> transforms;
    std::vector<int32_t> clip_aux;       // 2 per draw: enclosing BeginClip (-1), link (End for Begin, Begin for End)
    std::vector<int32_t> clip_stack;     // open BeginClip draw indices
    std::vector<uint8_t> clip_kind;      // 0 = clip, 1 = layer, 2 = layer merged with its clip shape (parallel to clip_stack)
    std::vector<float> clip_bb;          // 4 per open clip: device-space bounds outside which it shows nothing
    float next_clip_bb[4] = {-3.0e38f, -3.0e38f, 3.0e38f, 3.0e38f};   // bounds of the clip path just emitted (consumed by begin_clip)
    void set_next_clip_bounds(const float transform[6], const uint8_t* verbs, size_t n_verbs, const float* coords, size_t n_coords,
                              const uint8_t* verb_map = nullptr);
    uint32_t n_paths = 0, n_clips = 0, n_seg_tags = 0, n_implicit = 0;
    // Multi-GPU bands: fills and strokes of a scene.Encoding whose control points (plus the stroke's reach) stay outside the
    // pixel rows [cull_lo, cull_hi) of this device's band are dropped at ingest -- they own no tile there -- so that the
    // packed scene, its upload and every per-path stage shrink with the band. Clips and layers are always kept.
    bool cull = false; float cull_lo = 0, cull_hi = 0; uint32_t n_culled = 0;
    void set_cull_band(float lo, float hi) { cull = true; cull_lo = lo; cull_hi = hi; }
    bool outside_band(const float t[6], const float* c, size_t n_coords, float reach) const;
    float last_transform[6] = {0, 0, 0, 0, 0, 0};
    bool have_transform = false;
    // current path state
    bool in_path = false, has_move = false;
    float cur[2] = {0, 0}, start[2] = {0, 0};
    float path_t[6] = {1, 0, 0, 0, 1, 0};

    void clear(uint32_t w, uint32_t h);
    uint32_t n_draws() const { return (uint32_t)draw_tags.size(); }

    // path construction (coordinates in the space of `transform`; the device applies it)
    void begin_path(const float transform[6], bool even_odd);
    // Stroked path: the centre line travels to the device (which expands it, stroke.cuh); coordinates in the space of
    // `transform`, width in device units (scene/renderer.go:655-713 strokes the transformed points with the raw width).
    void stroke_path(const float transform[6], const uint8_t* verbs, size_t n_verbs, const float* coords, size_t n_coords,
                     const struct StrokeStyleHost& st, const uint8_t* verb_map = nullptr);   // verb_map: verbs[] byte -> GGCUDA_VERB_*
    void move_to(float x, float y);
    void line_to(float x, float y);
    void quad_to(float cx, float cy, float x, float y);
    void cubic_to(float c1x, float c1y, float c2x, float c2y, float x, float y);
    void close();
    void end_path();                      // auto-closes the open subpath, emits the Path marker
    void add_verbs(const uint8_t* verbs, uint32_t n_verbs, const double* coords, uint32_t n_coords);
    // a whole fill path (begin_path .. end_path) from verb bytes + float coordinates, without per-point calls
    void fill_verbs(const float transform[6], bool even_odd, const uint8_t* verbs, size_t n_verbs, const float* coords, size_t n_coords,
                    const uint8_t* verb_map = nullptr);

    void append_stroke(const StrokeSink& k);               // outline loops of an expanded stroke into the current path
    void draw_color(uint32_t rgba_premul);                 // DrawTagColor for the path just ended
    // Gradient brush for the path just ended (gg.LinearGradientBrush / RadialGradientBrush, gradient_linear.go,
    // gradient_radial.go, gradient.go). kind 0 linear: geom = x0, y0, x1, y1; kind 1 radial: geom = cx, cy, r0, r1 (device
    // space). stops: 5 floats each {offset, r, g, b, a} straight alpha. The table travels behind the packed scene (see pack()).
    void draw_gradient(int kind, const double geom[6], const double* stops, uint32_t n_stops, int extend);
    // TagFillRoundRect as the CPU renderer draws it (scene/renderer.go:986-1043, scene/shape.go:246-274): coverage from a signed
    // distance field with a 0.7 px smoothstep, not from the outline's area. Record kind 2 of the same table: cx, cy, half width,
    // half height, corner radius (device space), the premultiplied RGBA8 colour's bits. The path just ended only bins tiles.
    void draw_sdf_round_rect(float cx, float cy, float half_w, float half_h, float radius, uint32_t rgba_premul);
    // TagImage (scene/renderer.go:1093-1243 blitImageToTile): images are registered per frame (premultiplied RGBA8, the
    // bytes scene.Image.Data holds), drawn by index under their own affine. Record kind 3 of the same table: width, height,
    // word offset of the pixels (behind the stops), the INVERSE affine (float32, computed as the reference does), the
    // truncated device-space bounding box the reference iterates over. The path just ended (that box) only bins tiles.
    struct HostImage { uint32_t w, h; size_t off; };
    std::vector<HostImage> images;
    std::vector<uint32_t> image_words;
    int add_image(uint32_t w, uint32_t h, const uint8_t* premul_rgba);   // index of the image, -1 if the size is unusable
    bool draw_image(uint32_t index, const float t[6]);                   // false: no such image
    // gradient table words: per gradient a 16-word record {kind, extend, n_stops, stops offset, 0, 6 raw geometry floats, pad},
    // then all stops, sorted per gradient, 8 floats each {offset, r, g, b, a (straight sRGB), linear-light r, g, b}, then the
    // pixels of the frame's images (one word per pixel)
    std::vector<uint32_t> grad_recs; std::vector<float> grad_stops;
    uint32_t n_gradients = 0;
    size_t gradient_words() const { return grad_recs.size() + grad_stops.size() + image_words.size(); }
    void begin_clip(uint32_t blend_word, float alpha, uint8_t kind);   // DrawTagBeginClip for the path just ended
    void begin_layer(uint32_t blend_word, float alpha);   // PushLayer: clip rectangle carrying blend + alpha
    bool end_clip(uint8_t kind);                           // DrawTagEndClip (+ dummy path); false if nothing to pop
    void close_open_clips();

    // scene.Encoding ingest; returns 0 or a negative GGCUDA_ERR_* with msg set
    int add_encoding(const uint8_t* tags_in, size_t n_tags, const float* pd, size_t n_pd, const uint32_t* dd, size_t n_dd,
                     const float* tr, size_t n_tr, const double* brushes, size_t n_brushes, std::string* msg);

    // packed words + layout
    struct Layout {
        uint32_t n_tag_bytes, n_tag_words, n_draws, n_paths, n_clips;
        uint32_t path_tag_base, path_data_base, draw_tag_base, draw_data_base, transform_base, style_base, clip_aux_base, n_scene_words;
    };
    size_t packed_words() const;          // including the constant tail
    void pack(uint32_t* out, Layout* L, uint32_t band_tiles) const;
};

// premultiplied RGBA8 packing of a straight RGBA8 colour (scene_encode.go:162-168)
uint32_t gg_pack_color_straight(const uint8_t c[4]);
// clampU8 of a float colour channel (path_convert.go:131-140)
uint8_t gg_clamp_u8(double v);
// scene.BlendMode -> PTCL blend word: (mix << 8) | compose, Vello/peniko numbering
uint32_t gg_blend_word(uint32_t scene_blend_mode);

// Stroke outline of a device-space path as a polygon set to be filled NonZero
// (software.go:1145-1226 fills the expanded stroke with the paint's rule, NonZero).
struct StrokeStyleHost { double width, miter_limit; int cap, join; };
void gg_stroke_to_fill(const std::vector<uint8_t>& verbs, const std::vector<float>& coords, const StrokeStyleHost& st, StrokeSink* out);
