// b32_launch.h — host-side launch interface between b32_api.cu (C ABI, memory, streams) and
// b32_kernels.cu (kernels).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "b32_device.cuh"

namespace b32 {

// An instantiated CUDA graph of one enqueued frame (k_setup, k_fill_opaque[, k_fill_ordered]) whose
// kernel nodes are re-parameterised in place: when LaunchCtx.patch is set, the frame kernels' launchers write
// their arguments into the graph's nodes instead of launching, and the caller launches the graph once.
struct GraphPatch {
    cudaGraphExec_t exec = nullptr;
    int n = 0;
    cudaGraphNode_t node[8];
    void* func[8];
    bool used[8];
    cudaError_t err = cudaSuccess;
};

struct LaunchCtx {
    cudaStream_t stream;
    uint32_t sms;                 // SM count of the device: grids are sized in multiples of it
    uint64_t* launches;           // counter of this library's own kernel launches
    GraphPatch* patch = nullptr;  // non-null: patch graph nodes instead of launching
};

// cudaFuncSetAttribute for the kernels that need more than 48 KB of dynamic shared memory, on the current device.
// Returns a cudaError_t value.
int init_kernel_attributes();

void launch_transform(const LaunchCtx& L, const b32_vertex* verts, TVert* out, float* dbg_cam, const CallParams& p);
// tv == nullptr: vertices are transformed inside k_setup (fused path).  Also writes the tile masks (binning without bins,
// b32_device.cuh) and counts the ordered pass's entries per mask tile in ocount[] / CallState.obin_max.
// clear_n != 0: k_setup first clears the framebuffer (clear_n pixels to clear_color, depth to f32::MAX)
void launch_setup(const LaunchCtx& L, const b32_vertex* verts, const b32_face* faces, const TVert* tv, const TexDev* tex,
                  const LightDev* lights, SurfRec* recs, uint64_t* keys, BinHead* heads, uint4* masks, uint32_t* ocount,
                  WireTri* wire, CallState* st, uint32_t* zero_next, uint32_t zero_words,
                  uint32_t* clear_rgba, float* clear_z, uint32_t clear_n, uint32_t clear_color, const CallParams& p);
// pass 1: every tile reads its surface list out of its mask row and k_setup's per-face heads.  crowd: crowd_cap heads of
// scratch for tiles with more candidates than one window holds (they order their walk by depth through a slice of it;
// without one they take their windows in face order)
void launch_fill_opaque(const LaunchCtx& L, const SurfRec* recs, const uint4* masks, const BinHead* heads,
                        const TexDev* tex, const uint16_t* texels, const uint32_t* texmask, uint32_t* fb_rgba, float* fb_z,
                        CallState* st, uint32_t* sticky, BinHead* crowd, uint32_t crowd_cap, const CallParams& p);
// ordered pass (pass 2 / x-ray): every tile builds its draw-order entries from its mask row + keys[], sorts them and replays
// them in order.  scratch: scratch_cap (a power of two, or 0) entries per tile for tiles with more entries than fit shared memory.
// p.enq_ordered: launched right behind pass 1 without a host round trip (exits at once when there is nothing to replay).
void launch_fill_ordered(const LaunchCtx& L, const SurfRec* recs, const uint4* masks, BinHead* scratch, const uint64_t* keys,
                         const TexDev* tex, const uint16_t* texels, uint32_t* fb_rgba, float* fb_z,
                         CallState* st, uint32_t* sticky, const CallParams& p, uint32_t scratch_cap);
// wireframe phase: kind 1 = back-face edges (depth tested), 2 = front-face overlay edges
// table: table_size (a power of two >= 6 * nf) words of scratch for the first-occurrence edge de-duplication
void launch_wire(const LaunchCtx& L, const WireTri* wire, uint32_t kind, uint32_t color, bool depth_test, uint32_t* table, uint32_t table_size,
                 uint32_t* fb_rgba, const float* fb_z, CallState* st, const CallParams& p);
void launch_fb_clear(const LaunchCtx& L, uint32_t* rgba, float* z, uint32_t n, uint32_t color);
// Framebuffer::clear_gradient; top/bottom = r | g << 8 | b << 16 (| alpha byte << 24 in top)
void launch_fb_clear_gradient(const LaunchCtx& L, uint32_t* rgba, float* z, uint32_t w, uint32_t h, uint32_t top, uint32_t bottom);
// overlay lines (b32_overlay.cu): begin = last-overwrite claim + store + first proposals; then one launch per round.
// owner (= applied): one word per pixel; next: two proposal planes of one word per pixel, used alternately; wait: one
// word per line; flags[r] != 0 iff blended operations still wait after round r.
void launch_lines_begin(const LaunchCtx& L, const b32_line* lines, uint32_t n, uint32_t* owner, uint32_t* next, uint32_t* wait, uint32_t* flags,
                        uint32_t n_flags, uint32_t* fb_rgba, const float* fb_z, uint32_t w, uint32_t h, bool any_blended);
void launch_lines_round(const LaunchCtx& L, const b32_line* lines, uint32_t n, uint32_t* applied, uint32_t* next_cur, uint32_t* next_nxt,
                        uint32_t* wait, uint32_t* flags, uint32_t round, uint32_t* fb_rgba, const float* fb_z, uint32_t w, uint32_t h);
void launch_tex_expand(const LaunchCtx& L, const uint8_t* idx, const uint16_t* clut, uint32_t clut_len, uint32_t format, uint32_t n, uint16_t* out);

// skybox sphere pass (render.rs:81-139, :242-299): setup (+ tile masks) -> fill
void launch_sky(const LaunchCtx& L, const b32_sky_vertex* verts, const uint32_t* faces, SkyRec* recs, BinHead* heads, uint4* masks,
                uint32_t* fb_rgba, CallState* st, uint32_t* zero_next, uint32_t zero_words, const CallParams& p);

// render_asset_parts' vertex transform (scene.rs:141-160): out = in rotated about Y, translated
void launch_place(const LaunchCtx& L, const b32_vertex* in, b32_vertex* out, uint32_t nv, float cos_f, float sin_f, const float* world_pos);

// star field (render.rs:149-235): owner = one word per pixel of scratch; p.cam_pos must be zero
void launch_stars(const LaunchCtx& L, const b32_star* stars, uint32_t n, int32_t size, uint32_t* owner, uint32_t* fb_rgba, const CallParams& p);

// bit i of mask = texel i of the pool writes when its surface is black-keyed; n_words covers n_texels, zero padded
void launch_tex_mask(const LaunchCtx& L, const uint16_t* texels, uint32_t n_texels, uint32_t n_words, uint32_t* mask);

// RGB888 texel pool: the "texel writes" mask (tag != Erase) and, per texture, TexDev.blend = "has blended texels"
void launch_tex8_scan(const LaunchCtx& L, const uint32_t* texels, uint32_t n_texels, uint32_t n_words, uint32_t* mask,
                      TexDev* desc, uint32_t ntex, uint32_t max_texels);

}  // namespace b32
