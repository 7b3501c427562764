// b32_overlay.cu — framebuffer passes around the hot path, sm_100a.
//
//   k_fb_clear_gradient   Framebuffer::clear_gradient (render.rs:60-77)
//   k_lines_*             Framebuffer::draw_line / draw_line_blended / draw_line_alpha / draw_line_3d /
//                         draw_line_3d_overlay / draw_line_3d_alpha (render.rs:684-872) and the filled primitives
//                         draw_circle / draw_circle_alpha / draw_filled_rect / draw_thick_line (render.rs:631-682,
//                         :875-972) for a whole list, with the result of drawing them one after the other in list order.
//
// Order on the device.  A line never writes the z-buffer, so whether line i touches pixel p (bounds + depth test) is
// independent of every other line; only the colour at p depends on the order.  Two kinds of pixel operations exist:
//   overwrite  set_pixel / Erase: the result ignores what is there.  The last overwriting line at p (largest index)
//              is found with one atomicMax per touched pixel (k_lines_claim) and writes alone (k_lines_write).
//   blend      set_pixel_blended (Average/Add/Subtract/AddQuarter) and set_pixel_alpha read the pixel.  Those of a
//              pixel that come after its last overwrite are applied in index order, one per round (k_lines_round): every
//              pending line proposes its index with atomicMin, the smallest applies in the next round.  The number
//              of rounds is the deepest stack of blended lines over one pixel — 1 or 2 for overlays; lines with nothing
//              left to apply drop out, rounds with nothing pending return at once.
// Bresenham's walk has a closed form (walk_line), so every pixel of every line is an independent work item: one CTA
// per line, threads over its on-screen steps.
//
// Compiled with -fmad=false like the rest of the library: the depth interpolation is the reference's f32 sequence.

#include <cuda_runtime.h>
#include <stdint.h>

#include "b32_device.cuh"
#include "b32_raster.h"
#include "b32_launch.h"

namespace b32 {

namespace {

// ------------------------------------------------------------------------------------------------
// Framebuffer::clear_gradient (render.rs:60-77) with Color::lerp (types.rs:811-820)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lerp_channel(uint32_t a, uint32_t b, float inv_t, float t) {
    float v = (float)a * inv_t + (float)b * t;      // two roundings of the products, one of the sum (no fma)
    return (uint32_t)v;                             // in [0, 255]: `as u8` truncates
}

__global__ void k_fb_clear_gradient(uint32_t* __restrict__ rgba, float* __restrict__ z, uint32_t w, uint32_t h,
                                    uint32_t top, uint32_t bottom) {
    const uint32_t n = w * h;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t y = i / w;
        float t = h > 1 ? (float)y / (float)(h - 1) : 0.0f;         // :64
        t = fminf(fmaxf(t, 0.0f), 1.0f);                            // lerp's clamp (:812)
        float inv_t = 1.0f - t;
        uint32_t r = lerp_channel(top & 255u, bottom & 255u, inv_t, t);
        uint32_t g = lerp_channel((top >> 8) & 255u, (bottom >> 8) & 255u, inv_t, t);
        uint32_t b = lerp_channel((top >> 16) & 255u, (bottom >> 16) & 255u, inv_t, t);
        rgba[i] = r | (g << 8) | (b << 16) | (top & 0xFF000000u);  // blend tag (alpha byte) = top's (:818)
        z[i] = 3.40282347e+38f;
    }
}

// ------------------------------------------------------------------------------------------------
// lines
// ------------------------------------------------------------------------------------------------
constexpr uint32_t LINE_NONE = 0xFFFFFFFFu;
constexpr uint32_t LINE_BLOCK = 128;        // one CTA per line: the walks are chains of dependent L2 round trips, so a line
                                            // is spread over 128 lanes rather than 32

__device__ __forceinline__ bool line_overwrites(const b32_line& l) {
    if (l.kind == B32_LINE_2D) return l.mode == B32_BLEND_OPAQUE || l.mode == B32_BLEND_ERASE;
    return l.kind != B32_LINE_2D_ALPHA && l.kind != B32_LINE_3D_ALPHA && l.kind != B32_LINE_CIRCLE_ALPHA;
}

// what an overwriting line stores: Color::to_bytes (types.rs:829-832); set_pixel_blended with mode Erase stores
// Color::TRANSPARENT (types.rs:920-923) = (0, 0, 0, Erase) -> bytes 0, 0, 0, 0
__device__ __forceinline__ uint32_t line_store_value(const b32_line& l) {
    if (l.kind == B32_LINE_2D && l.mode == B32_BLEND_ERASE) return 0u;
    uint32_t a = l.blend == B32_BLEND_ERASE ? 0u : 255u;
    return (uint32_t)l.r | ((uint32_t)l.g << 8) | ((uint32_t)l.b << 16) | (a << 24);
}

// set_pixel_blended (render.rs:313-333, Color::blend_with types.rs:886-930) / set_pixel_alpha (render.rs:646-667)
__device__ __forceinline__ uint32_t line_blend_value(const b32_line& l, uint32_t back) {
    uint32_t br = back & 255u, bg = (back >> 8) & 255u, bb = (back >> 16) & 255u;
    uint32_t r, g, b;
    if (l.kind == B32_LINE_2D) {
        r = blend8(l.r, br, l.mode); g = blend8(l.g, bg, l.mode); b = blend8(l.b, bb, l.mode);
    } else {
        uint32_t a = l.alpha, ia = 255u - a;
        r = (l.r * a + br * ia) / 255u; g = (l.g * a + bg * ia) / 255u; b = (l.b * a + bb * ia) / 255u;
    }
    return r | (g << 8) | (b << 16) | 0xFF000000u;
}

// The walk shared by all six reference functions, in closed form.  With a = |x1 - x0|, b = |y1 - y0| the reference loop
//     e2 = 2 * err;  if e2 >= -b { err -= b; x += sx }  if e2 <= a { err += a; y += sy }      (err starts at a - b)
// moves its major axis in every iteration (a >= b: 2 * err >= a - 2b + 1 >= -b holds throughout; b > a: 2 * err <= a)
// and its minor axis in iteration k -> k + 1 iff 2 * major * n_k + major <= 2 * minor * (k + 1), i.e.
//     n_k = floor((2 * minor * k + major) / (2 * major))
// and the depth interpolation's `step` counter is k (exactly one of its two increments fires per iteration).  So pixel k
// of a line is computed independently: one CTA per line, threads stride over the part of the walk whose MAJOR coordinate
// is on screen.  visit(pixel index) is called for the on-screen pixels that pass the depth test.
// tests/test_gpu_parity.py::test_draw_lines_every_slope checks the closed form against the reference loop for every
// slope of a 41 x 41 neighbourhood; i32/u64 arithmetic cannot overflow with |coordinate| <= 2^20 (checked by the host).
template <class Visit>
__device__ __forceinline__ void walk_line(const b32_line& l, int32_t W, int32_t H, const float* __restrict__ fb_z,
                                          uint32_t lane, Visit visit) {
    const uint32_t a = (uint32_t)abs(l.x1 - l.x0), b = (uint32_t)abs(l.y1 - l.y0);
    const int32_t sx = l.x0 < l.x1 ? 1 : -1, sy = l.y0 < l.y1 ? 1 : -1;
    const bool xmajor = a >= b;
    const uint32_t major = xmajor ? a : b, minor = xmajor ? b : a;
    // k range with the major coordinate inside the screen
    const int32_t m0 = xmajor ? l.x0 : l.y0, ms = xmajor ? sx : sy, mlim = xmajor ? W : H;
    int64_t lo = ms > 0 ? -(int64_t)m0 : (int64_t)m0 - (mlim - 1);
    int64_t hi = ms > 0 ? (int64_t)(mlim - 1) - m0 : (int64_t)m0;
    if (lo < 0) lo = 0;
    if (hi > (int64_t)major) hi = major;
    if (lo > hi) return;
    const bool depth = l.kind == B32_LINE_3D || l.kind == B32_LINE_3D_OVERLAY || l.kind == B32_LINE_3D_ALPHA;
    const bool allow_equal = l.kind != B32_LINE_3D;
    float z0 = l.z0, z1 = l.z1;
    if (l.kind == B32_LINE_3D_ALPHA) { z0 = z0 * 0.995f; z1 = z1 * 0.995f; }       // DEPTH_BIAS (:826-828)
    const float total_steps = (float)max(major, 1u);                               // :776 / :839
    const float dz = z1 - z0;
    for (uint32_t k = (uint32_t)lo + lane; k <= (uint32_t)hi; k += LINE_BLOCK) {
        uint32_t n = 0;
        if (minor) {
            uint64_t num = 2ull * minor * k + major;
            n = (num >> 32) ? (uint32_t)(num / (2ull * major)) : (uint32_t)num / (2u * major);
        }
        const int32_t x = xmajor ? l.x0 + sx * (int32_t)k : l.x0 + sx * (int32_t)n;
        const int32_t y = xmajor ? l.y0 + sy * (int32_t)n : l.y0 + sy * (int32_t)k;
        if (x < 0 || x >= W || y < 0 || y >= H) continue;
        const uint32_t idx = (uint32_t)y * (uint32_t)W + (uint32_t)x;
        if (depth) {
            float t = (float)k / total_steps;                                       // `step` is an exact integer < 2^24
            float z = z0 + t * dz;
            float zb = fb_z[idx];
            if (!(allow_equal ? z <= zb : z < zb)) continue;
        }
        visit(idx);
    }
}

// The filled primitives: every pixel of the clamped bounding box is one work item.
//   draw_circle / draw_circle_alpha (render.rs:631-644, :670-682), draw_filled_rect (:954-972), draw_thick_line (:875-938)
template <class Visit>
__device__ __forceinline__ void walk_area(const b32_line& l, int32_t W, int32_t H, uint32_t lane, Visit visit) {
    int32_t bx0, bx1, by0, by1;                         // inclusive, clamped to the screen
    float cx[4] = {0, 0, 0, 0}, cy[4] = {0, 0, 0, 0};   // thick line: the quad's corners
    if (l.kind == B32_LINE_CIRCLE || l.kind == B32_LINE_CIRCLE_ALPHA) {
        const int32_t r = l.x1;
        by0 = max(l.y0 - r, 0); by1 = min(l.y0 + r, H - 1);
        bx0 = max(l.x0 - r, 0); bx1 = min(l.x0 + r, W - 1);
    } else if (l.kind == B32_LINE_FILLED_RECT) {
        bx0 = max(min(l.x0, l.x1), 0); bx1 = min(max(l.x0, l.x1), W - 1);
        by0 = max(min(l.y0, l.y1), 0); by1 = min(max(l.y0, l.y1), H - 1);
    } else {                                            // B32_LINE_THICK with thickness > 1
        const float dx = (float)(l.x1 - l.x0), dy = (float)(l.y1 - l.y0);
        const float len = sqrtf(dx * dx + dy * dy);                                  // :884
        if (len < 0.001f) return;
        const float half = l.z0 * 0.5f;                                              // thickness as f32 * 0.5
        const float px = -dy / len * half, py = dx / len * half;
        const float fx0 = (float)l.x0, fy0 = (float)l.y0, fx1 = (float)l.x1, fy1 = (float)l.y1;
        cx[0] = fx0 + px; cy[0] = fy0 + py; cx[1] = fx0 - px; cy[1] = fy0 - py;      // :894-899
        cx[2] = fx1 - px; cy[2] = fy1 - py; cx[3] = fx1 + px; cy[3] = fy1 + py;
        bx0 = max(f2i32(fminf(fminf(cx[0], cx[1]), fminf(cx[2], cx[3]))), 0);         // :902-911 (`as i32` truncates)
        bx1 = min(f2i32(fmaxf(fmaxf(cx[0], cx[1]), fmaxf(cx[2], cx[3]))), W - 1);
        by0 = max(f2i32(fminf(fminf(cy[0], cy[1]), fminf(cy[2], cy[3]))), 0);
        by1 = min(f2i32(fmaxf(fmaxf(cy[0], cy[1]), fmaxf(cy[2], cy[3]))), H - 1);
    }
    if (bx0 > bx1 || by0 > by1) return;
    const uint32_t bw = (uint32_t)(bx1 - bx0 + 1), n = bw * (uint32_t)(by1 - by0 + 1);
    for (uint32_t k = lane; k < n; k += LINE_BLOCK) {
        const int32_t x = bx0 + (int32_t)(k % bw), y = by0 + (int32_t)(k / bw);
        if (l.kind == B32_LINE_CIRCLE || l.kind == B32_LINE_CIRCLE_ALPHA) {
            const int32_t dx = x - l.x0, dy = y - l.y0;
            if (dx * dx + dy * dy > l.x1 * l.x1) continue;                           // :637-639
        } else if (l.kind == B32_LINE_THICK) {
            const float ppx = (float)x + 0.5f, ppy = (float)y + 0.5f;                // :924
            bool inside = true;
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = (i + 1) & 3;
                const float cross = (cx[j] - cx[i]) * (ppy - cy[i]) - (cy[j] - cy[i]) * (ppx - cx[i]);   // :929
                if (cross < 0.0f) inside = false;
            }
            if (!inside) continue;
        }
        visit((uint32_t)y * (uint32_t)W + (uint32_t)x);
    }
}

template <class Visit>
__device__ __forceinline__ void walk_prim(const b32_line& l, int32_t W, int32_t H, const float* __restrict__ fb_z, uint32_t lane, Visit visit) {
    const bool area = l.kind == B32_LINE_CIRCLE || l.kind == B32_LINE_CIRCLE_ALPHA || l.kind == B32_LINE_FILLED_RECT ||
                      (l.kind == B32_LINE_THICK && l.z0 > 1.0f);                     // thickness <= 1: draw_line (:876-879)
    if (area) walk_area(l, W, H, lane, visit);
    else walk_line(l, W, H, fb_z, lane, visit);
}

// last overwriting line of every pixel
__global__ void __launch_bounds__(LINE_BLOCK)
k_lines_claim(const b32_line* __restrict__ lines, uint32_t n, uint32_t* __restrict__ owner, const float* __restrict__ fb_z,
              int32_t W, int32_t H) {
    const uint32_t i = blockIdx.x, lane = threadIdx.x;
    if (i >= n) return;
    const b32_line l = lines[i];
    if (!line_overwrites(l)) return;
    walk_prim(l, W, H, fb_z, lane, [&](uint32_t idx) { atomicMax(&owner[idx], i + 1); });
}

// overwriting lines: the last one of each pixel stores.  Blended lines: first proposal round.
__global__ void __launch_bounds__(LINE_BLOCK)
k_lines_write(const b32_line* __restrict__ lines, uint32_t n, const uint32_t* __restrict__ owner, uint32_t* __restrict__ next,
              uint32_t* __restrict__ fb_rgba, const float* __restrict__ fb_z, int32_t W, int32_t H) {
    const uint32_t i = blockIdx.x, lane = threadIdx.x;
    if (i >= n) return;
    const b32_line l = lines[i];
    if (line_overwrites(l)) {
        const uint32_t v = line_store_value(l);
        walk_prim(l, W, H, fb_z, lane, [&](uint32_t idx) { if (owner[idx] == i + 1) fb_rgba[idx] = v; });
    } else {
        walk_prim(l, W, H, fb_z, lane, [&](uint32_t idx) { if (i + 1 > owner[idx]) atomicMin(&next[idx], i + 1); });
    }
}

// One round: `applied[p]` = index + 1 of the last operation applied to pixel p (starts as its last overwrite).  The line
// whose index is the pixel's proposal in `next_cur` applies; every other waiting line proposes itself for the next round
// in `next_nxt` (the two planes alternate, the applier hands its entry back empty).  `wait[i]` != 0 while line i has
// operations left (all lines start waiting); flags[r] != 0 iff any line waits after round r.
__global__ void __launch_bounds__(LINE_BLOCK)
k_lines_round(const b32_line* __restrict__ lines, uint32_t n, uint32_t* __restrict__ applied, uint32_t* __restrict__ next_cur,
              uint32_t* __restrict__ next_nxt, uint32_t* __restrict__ fb_rgba, const float* __restrict__ fb_z, int32_t W, int32_t H,
              uint32_t* __restrict__ wait, uint32_t* __restrict__ flags, uint32_t round) {
    if (round > 0 && flags[round - 1] == 0) return;
    const uint32_t i = blockIdx.x, lane = threadIdx.x;
    if (i >= n || wait[i] == 0) return;
    const b32_line l = lines[i];
    if (line_overwrites(l)) { if (lane == 0) wait[i] = 0; return; }
    bool waiting = false;
    walk_prim(l, W, H, fb_z, lane, [&](uint32_t idx) {
        // a line visits a pixel once (the walk moves in every iteration), so `applied` is never this line's own write;
        // a concurrent applier's index is below every waiting line's, so a stale read decides the same
        if (i + 1 <= applied[idx]) return;
        if (next_cur[idx] == i + 1) {                   // no other thread matches this pixel's proposal: plain accesses are safe
            fb_rgba[idx] = line_blend_value(l, fb_rgba[idx]);
            applied[idx] = i + 1;
            next_cur[idx] = LINE_NONE;
        } else {
            waiting = true;
            atomicMin(&next_nxt[idx], i + 1);
        }
    });
    waiting = __syncthreads_or(waiting);
    if (lane == 0) {
        wait[i] = waiting;
        if (waiting) flags[round] = 1;
    }
}

inline uint32_t blocks_for(uint32_t n, uint32_t block) { return (n + block - 1) / block; }

}  // namespace

void launch_fb_clear_gradient(const LaunchCtx& L, uint32_t* rgba, float* z, uint32_t w, uint32_t h, uint32_t top, uint32_t bottom) {
    if (w == 0 || h == 0) return;
    uint32_t n = w * h, g = blocks_for(n, 256), cap = L.sms * 8;
    k_fb_clear_gradient<<<g > cap ? cap : g, 256, 0, L.stream>>>(rgba, z, w, h, top, bottom);
    ++*L.launches;
}

static inline uint32_t line_grid(uint32_t n) { return n; }

void launch_lines_begin(const LaunchCtx& L, const b32_line* lines, uint32_t n, uint32_t* owner, uint32_t* next, uint32_t* wait, uint32_t* flags,
                        uint32_t n_flags, uint32_t* fb_rgba, const float* fb_z, uint32_t w, uint32_t h, bool any_blended) {
    cudaMemsetAsync(owner, 0, (size_t)w * h * 4, L.stream);
    k_lines_claim<<<line_grid(n), LINE_BLOCK, 0, L.stream>>>(lines, n, owner, fb_z, (int32_t)w, (int32_t)h);
    if (any_blended) {
        cudaMemsetAsync(next, 0xFF, (size_t)w * h * 8, L.stream);           // both proposal planes
        cudaMemsetAsync(wait, 1, (size_t)n * 4, L.stream);
        cudaMemsetAsync(flags, 0, (size_t)n_flags * 4, L.stream);
    }
    k_lines_write<<<line_grid(n), LINE_BLOCK, 0, L.stream>>>(lines, n, owner, next, fb_rgba, fb_z, (int32_t)w, (int32_t)h);
    *L.launches += 2;
}

void launch_lines_round(const LaunchCtx& L, const b32_line* lines, uint32_t n, uint32_t* applied, uint32_t* next_cur, uint32_t* next_nxt,
                        uint32_t* wait, uint32_t* flags, uint32_t round, uint32_t* fb_rgba, const float* fb_z, uint32_t w, uint32_t h) {
    k_lines_round<<<line_grid(n), LINE_BLOCK, 0, L.stream>>>(lines, n, applied, next_cur, next_nxt, fb_rgba, fb_z, (int32_t)w, (int32_t)h, wait, flags, round);
    ++*L.launches;
}

}  // namespace b32
