// b32_device.cuh — device-side records and exact-arithmetic helpers shared by the kernels.
//
// Bit-exactness contract: every f32 expression is evaluated left to right with one
// rounding per operator and never fused.  The translation unit is compiled with
//   -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
// and the helpers below restate the Rust cast / min / max / rem_euclid semantics.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "b32_raster.h"

namespace b32 {

constexpr int TILE_W = 16;            // screen tile owned by one CTA of k_fill
constexpr int TILE_H = 16;
constexpr int FILL_THREADS = TILE_W * TILE_H;
constexpr float NEAR_PLANE = 0.1f;    // math.rs:155
constexpr uint32_t WIRE_MAX_STEPS = 1u << 24;   // longest Bresenham walk k_wire performs (the reference walks every step, on or off screen)
constexpr int OP_SORT_MAX_ENTRIES = 1024;  // k_fill_opaque orders a tile's candidates in shared memory this many at a time (a window)
constexpr int ORD_SORT_MAX = 2048;         // draw-order entries k_fill_ordered sorts in shared memory (32 KB); tiles with more use a slice of a global scratch
constexpr int SETUP_GROUP = 128;           // faces per k_setup CTA pass = bits of one tile-mask entry (uint4)
constexpr uint32_t MASK_TILES_MAX = 2048;  // mask tiles per frame (k_setup keeps one uint4 per mask tile in shared memory: 32 KB)
constexpr int OP_MASK_SMEM_WORDS = 2048;   // k_fill_opaque keeps the "texel writes" mask in shared memory up to 65536 texels (8 KB)

// ---- per-vertex output of k_transform (render.rs:2321-2360) ------------------------------------
// x,y,z = projected[i]; w = cam_space_positions[i].z (the only camera-space value the path reads:
// near-plane test :2381-2385 and fog :2419-2436; cam x/y only feed the dead `normal`).
typedef float4 TVert;

// ---- one drawable surface, ready for the fill (struct Surface render.rs:975-1000 after the
//      per-triangle preamble of rasterize_triangle_15, render.rs:1450-1527) -----------------------
struct __align__(16) SurfRec {
    float a0, b0, a1, b1;                 // edge-function steps, render.rs:1507-1510
    float w0s, w1s, inv_area;             // row-start values at (min_x,min_y) :1517-1518, 1/area :1504
    uint32_t flags;                       // SF_* below | editor_alpha << 8 | tex_id << 16
    uint32_t bbox_x, bbox_y;              // min | max << 16 (max exclusive), render.rs:1455-1458
    float iz1, iz2, iz3;                  // 1/v.z, render.rs:1546-1548 (pure function of the vertices)
    float u1, v1, u2, v2, u3, v3;         // Surface.uv1..3
    uint32_t vc1, vc2, vc3;               // Surface.vc1..3 as r | g<<8 | b<<16
    float sh[9];                          // Gouraud: 3 vertices x rgb (:1475-1483); Flat: sh[0..2] (:1466-1472)
    uint32_t _pad;
};
static_assert(sizeof(SurfRec) == 128, "SurfRec must be 128 bytes");

// The first 80 bytes of a SurfRec: everything the visibility walk of k_fill_opaque needs (edge functions, bbox, depth,
// texture coordinates).  Same field names, so the fragment helpers are templates over the record type.  Staged at an
// 80-byte stride, lane-indexed reads hit 8 different banks (a 128-byte stride puts all 32 lanes on one).
struct __align__(16) SurfHot {
    float a0, b0, a1, b1;
    float w0s, w1s, inv_area;
    uint32_t flags;
    uint32_t bbox_x, bbox_y;
    float iz1, iz2, iz3;
    float u1, v1, u2, v2, u3, v3;
    uint32_t vc1;
};
static_assert(sizeof(SurfHot) == 80, "SurfHot must be the first 80 bytes of SurfRec");

enum : uint32_t {
    SF_BLEND_MASK   = 0x7,        // blend_mode the fill uses: texture's if textured else face's (:1450-1452)
    SF_BLACK_TR     = 1u << 3,    // Face.black_transparent
    SF_DITHER       = 1u << 4,    // needs_dither, render.rs:1487-1492
    SF_TEXTURED     = 1u << 5,    // texture.is_some()
    SF_FAST_EDGE    = 1u << 6,    // edge stepping provably exact => closed form allowed
    SF_TRANSPARENT  = 1u << 7,    // has_transparency: drawn in pass 2 with skip_z_write (:2561-2569)
};

// Binning without bins.  k_setup handles SETUP_GROUP consecutive faces per CTA pass ("group") and writes, for every
// mask tile of the screen, ONE uint4 = 128 bits: bit i set iff face group*128 + i is drawn and its bounding box touches the
// tile.  Layout masks[mask_tile][group].  A fill CTA reads its tile's row (n_groups x 16 contiguous bytes), and the set
// bits ARE its surface list — no atomics, no per-tile capacity, no overflow, deterministic.  A mask tile is
// (16 << mshift) pixels square: mshift grows with the frame so that the row count stays <= MASK_TILES_MAX and the
// table <= MASK_BYTES_MAX; with mshift > 0 a fill CTA drops the candidates whose bounding box misses its own 16x16 tile.
constexpr size_t MASK_BYTES_MAX = (size_t)64 << 20;

// 16-byte per-face record next to the SurfRec: what a warp needs to reject a surface without
// touching its 128-byte record.
struct __align__(16) BinHead {
    uint32_t bbox_x, bbox_y;              // as in SurfRec
    uint32_t key;                         // depth_key_desc(center_z) (painter's priority), 0 in z-buffer mode
    uint32_t face;                        // Surface.face_idx == index of the SurfRec
};
static_assert(sizeof(BinHead) == 16, "BinHead must be 16 bytes");

// One drawable triangle of the skybox sphere pass (rasterize_skybox_triangle's arguments, render.rs:242-250)
struct __align__(16) SkyRec {
    float p0x, p0y, p1x, p1y, p2x, p2y, inv_denom;
    uint32_t c0, c1, c2;                  // vertex colours r | g<<8 | b<<16
    uint32_t _pad0, _pad1;
};
static_assert(sizeof(SkyRec) == 48, "SkyRec must be 48 bytes");

// A triangle collected for the wireframe phase (render.rs:2448-2450, :2509-2511): projected x, y, z of
// the ORIGINAL (unswapped) v1, v2, v3.  kind: 0 none, 1 back face, 2 front face.
struct WireTri { float x[3], y[3], z[3]; uint32_t kind; };

struct TexDev { uint32_t off, w, h, blend; };   // texel pool offset (u16 units), size, Texture15.blend_mode

struct LightDev {                     // b32_light without padding surprises
    uint32_t type; float px, py, pz, dx, dy, dz, radius, angle, intensity, cr, cg, cb; uint32_t enabled;
};

// device-side counters / flags of one render call
struct CallState {
    uint32_t n_opaque, n_transp;          // drawn surfaces per pass
    uint32_t nan_opaque, nan_transp;      // a NaN sort key was seen in the pass
    uint32_t oob;                         // a face index >= nv was seen
    uint32_t obin_overflow;               // ordered pass: a tile has more draw-order entries than its scratch slice (tile skipped, host grows + retries)
    uint32_t obin_max;                    // ordered pass: largest such count seen
    uint32_t wire_too_long;               // wireframe phase: an edge longer than WIRE_MAX_STEPS was skipped (host reports B32_ERR_UNSUPPORTED)
    uint32_t crowd_used;                  // pass 1: entries of the crowded-tile scratch handed out so far (one atomic per crowded tile)
    uint32_t n_big_stepped;               // pass-1 surfaces covering >= 1/64 of the framebuffer whose edge values are stepped (no SF_FAST_EDGE) in a fixed-point call
    uint32_t done[3];                     // blocking calls: CTAs of k_setup / k_fill_opaque / k_fill_ordered that have finished (HostStatus)
};

// Blocking calls (the reference's calling convention: render_mesh_15 returns timings and the drawn count) do not copy
// the CallState back and do not bracket the kernels with events: each kernel's last CTA publishes the kernel's end time
// — k_setup's also the counters — in host-mapped memory, and the host spins on `seq`.  That keeps the programmatic launch
// chain intact, saves the copy engine's round trip, and lets the host launch the ordered pass while pass 1 still runs.
enum { HS_SETUP = 0, HS_FILL = 1, HS_ORDERED = 2 };
struct KernelStamp {
    unsigned long long t0, t1;            // %globaltimer (ns): first CTA past its wait for the previous kernel, last CTA done
    uint32_t seq;                         // = CallParams.host_seq once t1 (and the state) are valid
    uint32_t _pad;
};
struct HostStatus {
    CallState state;                      // as k_setup left it
    uint32_t _pad[3];
    KernelStamp stamp[3];
};

// The reference panics (and draws nothing) on an out-of-range vertex index, or when a NaN key is
// compared by the sort, i.e. in any sorted slice of length >= 2 (render.rs:2531).
// render_mesh (RGB888) sorts ONE list and only in painter's mode (render.rs:2155-2162): k_setup then reports
// every NaN key in nan_opaque and the list length is n_opaque + n_transp.
__device__ __forceinline__ bool call_aborts(const CallState& st, bool use_zbuffer, bool rgb888) {
    if (st.oob) return true;
    if (rgb888) return !use_zbuffer && st.nan_opaque && st.n_opaque + st.n_transp >= 2;
    if (st.nan_transp && st.n_transp >= 2) return true;
    if (!use_zbuffer && st.nan_opaque && st.n_opaque >= 2) return true;
    return false;
}

// kernel parameters of one render call (passed by value => constant bank)
struct CallParams {
    float cam_pos[3], bx[3], by[3], bz[3];
    int32_t fcam_pos[3], fbx[3], fby[3], fbz[3];      // Fixed32::from_f32 of the above (fixed.rs:370-373)
    int32_t viewport_scale, half_w, half_h;           // fixed.rs:398-400
    uint32_t width, height, tiles_x, tiles_y;
    uint32_t nv, nf, ntex, n_lights;
    uint32_t n_groups;                                // ceil(nf / SETUP_GROUP): entries per row of the tile-mask table
    uint32_t mshift, mtiles_x, mtiles_y;              // mask tile = (16 << mshift) px square; mask tiles per row / column
    uint32_t mask_smem_words;                         // words of the "texel writes" mask to stage in shared memory (0: read it from global)
    uint8_t affine_textures, use_zbuffer, shading, backface_cull, dithering, use_fixed_point, xray_mode, ortho;
    uint8_t fog_enabled, fog_r, fog_g, fog_b, fog_blend, async_call, wire_back, wire_front;   // wire_*: render.rs:2576, :2606
    uint8_t rgb888;                                   // 1: render_mesh / rasterize_triangle (render.rs:1971-2259, 1202-1433)
    uint8_t enq_ordered;                              // 1: the ordered pass is enqueued behind pass 1 without a host round trip (k_fill_ordered exits early when it has nothing to do)
    uint8_t vwords;                                   // words per vertex record: 9 (b32_vertex) or 6 (b32_vertex_nn: no normal; only with shading None)
    uint8_t faces_implicit;                           // 1: `faces` holds one flags word per face; face i uses vertices 3i, 3i+1, 3i+2.  2: no face buffer at all, every face has `uniform_flags`
    float ambient, ortho_zoom, ortho_cx, ortho_cy;
    float fog_start, fog_falloff, fog_cull;
    uint8_t prefer_prefix;                            // host's hint: one of the last calls on this context drew large stepped surfaces, run the shared-prefix fill
    uint8_t has_spot;                                 // an enabled Spot light is in the list and shading is on: k_setup<true> (the acos path) runs
    uint32_t uniform_flags;                           // faces_implicit == 2: the flags word of every face
    uint32_t call_seq;                                // this context's call counter; the first large stepped surface of a fixed-point call stores it in *stepped_seq_host
    uint32_t* stepped_seq_host;                       // host-mapped word (never null)
    uint32_t host_seq;                                // blocking calls: the value the kernels publish in HostStatus when done
    HostStatus* host;                                 // null for enqueue-only calls
};

// Calls whose surfaces cannot have exact-integer edge values (float or ortho projection: SF_FAST_EDGE is never set) run the
// pass-1 fill with the shared edge prefix (k_fill_opaque<.., PRE = true>); also part of an enqueued frame's graph key.
__host__ __device__ inline bool fill_uses_edge_prefix(const CallParams& p) { return !p.use_fixed_point || p.ortho; }

// ---- Rust scalar semantics -------------------------------------------------------------------
// `f as i32` / `f as u32-ish`: cvt.rzi saturates and maps NaN to 0, exactly like Rust's `as`.
__device__ __forceinline__ int32_t f2i32(float f) { return __float2int_rz(f); }
__device__ __forceinline__ uint32_t f2u32sat(float f) { return __float2uint_rz(f); }   // usize, clipped to u32
__device__ __forceinline__ uint32_t f2u8(float f) { return min(__float2uint_rz(f), 255u); }
// f32::min/max = IEEE minNum/maxNum = fminf/fmaxf.  f32::clamp keeps NaN:
__device__ __forceinline__ float rclamp(float x, float lo, float hi) { if (x < lo) x = lo; if (x > hi) x = hi; return x; }
// f32::rem_euclid(1.0): r = fmod(u,1); if r < 0 { r + 1 }.  fmod(u,1) = u - trunc(u) exactly.
__device__ __forceinline__ float rem_euclid1(float u) {
    float r = u - truncf(u);                 // inf - inf = NaN, NaN stays NaN: same as fmodf
    return r < 0.0f ? r + 1.0f : r;
}

// ---- fixed.rs -------------------------------------------------------------------------------
__device__ __forceinline__ int32_t fx_from_f32(float f) { return f2i32(f * 4096.0f); }                 // :125-127
__device__ __forceinline__ int32_t fx_mul(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * (int64_t)b) >> 12); }  // :161-165
__device__ __forceinline__ int32_t fx_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int32_t fx_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

// UNR_TABLE[i] (fixed.rs:18-31): 257 bytes built at compile time from the reference's formula and read through the
// read-only path (neighbouring lanes hit the same three cache lines); a per-thread u32 division cost ~20 instructions
// per projected vertex.
struct UnrTable {
    uint8_t v[260];
    constexpr UnrTable() : v() {
        for (uint32_t i = 0; i <= 0x100u; ++i) {
            int32_t e = (int32_t)((0x40000u / (i + 0x100u) + 1u) >> 1) - 0x101;
            v[i] = (uint8_t)(e > 0 ? e : 0);
        }
    }
};
__device__ constexpr UnrTable g_unr_table{};
__device__ __forceinline__ uint32_t unr_entry(uint32_t i) { return __ldg(&g_unr_table.v[i]); }

// UNR reciprocal of a non-zero divisor (fixed.rs:183-205): returns nr2 and the final shift.  Every intermediate of the
// reference's u64 arithmetic fits 32 bits (d16 <= 0xFFFF, u <= 0x200: d16 * u < 2^25; nr1 < 2^18: nr1 * u < 2^27), so
// 32-bit arithmetic gives the same values.
__device__ __forceinline__ void unr_recip(int32_t divisor, uint32_t* nr2, uint32_t* shift) {
    uint32_t den = divisor < 0 ? 0u - (uint32_t)divisor : (uint32_t)divisor;
    uint32_t z = __clz(den);
    uint32_t d16 = (den << z) >> 16;                              // 0x8000..0xFFFF
    uint32_t idx = min((d16 - 0x7FC0u) >> 7, 256u);
    uint32_t u = unr_entry(idx) + 0x101u;
    uint32_t nr1 = (0x2000080u - d16 * u) >> 8;
    *nr2 = (0x80u + nr1 * u) >> 8;
    *shift = 36u - z;
}
// fixed.rs:207-230 given the reciprocal: |num| * nr2 is one 32x32 -> 64-bit multiply
__device__ __forceinline__ int32_t unr_apply(int32_t num, int32_t divisor, uint32_t nr2, uint32_t shift) {
    bool neg = (num < 0) != (divisor < 0);
    uint32_t n = num < 0 ? 0u - (uint32_t)num : (uint32_t)num;
    uint64_t raw = (uint64_t)n * nr2;
    uint64_t mag = (raw + (1ull << (shift - 1))) >> shift;       // shift in 5..36
    int32_t c = (int32_t)min((unsigned long long)mag, 0x7FFFFFFFull);
    return neg ? -c : c;
}

// ---- colour helpers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expand5(uint32_t v) { return ((v << 3) | (v >> 2)) & 0xFF; }   // render.rs:1161-1163

// blend_rgb555 for one channel (render.rs:1093-1145): 8-bit in, 5-bit math, result << 3
__device__ __forceinline__ uint32_t blend5(uint32_t f8, uint32_t b8, uint32_t mode) {
    uint32_t f5 = f8 >> 3, b5 = b8 >> 3, r;
    switch (mode) {
        case B32_BLEND_OPAQUE:      r = f5; break;
        case B32_BLEND_AVERAGE:     r = min((b5 + f5) >> 1, 31u); break;
        case B32_BLEND_ADD:         r = min(b5 + f5, 31u); break;
        case B32_BLEND_SUBTRACT:    r = b5 > f5 ? b5 - f5 : 0u; break;
        case B32_BLEND_ADD_QUARTER: r = min(b5 + (f5 >> 2), 31u); break;
        default:                    r = b5; break;     // Erase
    }
    return r << 3;
}

// Color::blend_with for one channel (types.rs:886-930): 8-bit math; Erase never reaches a writer
__device__ __forceinline__ uint32_t blend8(uint32_t f8, uint32_t b8, uint32_t mode) {
    switch (mode) {
        case B32_BLEND_AVERAGE:     return (b8 + f8) >> 1;
        case B32_BLEND_ADD:         return min(b8 + f8, 255u);
        case B32_BLEND_SUBTRACT:    return b8 > f8 ? b8 - f8 : 0u;
        case B32_BLEND_ADD_QUARTER: return min(b8 + (f8 >> 2), 255u);
        default:                    return f8;         // Opaque
    }
}

// order-preserving key for a stable ASCENDING radix sort that yields back-to-front order:
// descending center_z, +0 == -0 (render.rs:2527-2532)
__device__ __forceinline__ uint32_t depth_key_desc(float z) {
    uint32_t b = __float_as_uint(z);
    if (z == 0.0f) b = 0;
    uint32_t asc = b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
    return ~asc;
}

}  // namespace b32
