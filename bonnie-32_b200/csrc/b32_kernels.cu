// b32_kernels.cu — sm_100a kernels of the BONNIE-32 rasterizer hot path.
//
//   k_setup        render.rs:2321-2360 + fixed.rs:362-441 (vertex transform + snap) fused with render.rs:2373-2513 (cull /
//                  surface build / fog), :1450-1527 (triangle setup), :1013-1071 (lighting) and the sort key (:2527-2532):
//                  one face per thread, one 128-byte SurfRec + one 16-byte bin head per drawn face
//   k_bin_opaque   bin heads -> per-tile bins (any order), block-aggregated atomics; also bins the draw-order keys of pass 2
//   k_fill_opaque  render.rs:1530-1713 for pass 1 (opaque surfaces), order-free (see below); <true> = render_mesh (RGB888)
//   k_fill_ordered pass 2 (semi-transparent surfaces, back to front), x-ray mode and RGB888 calls that can blend:
//                  per-tile sort by the unique draw-order key, then strict in-order replay
//   k_wire_dedup + k_wire   the wireframe phase (render.rs:2574-2635): first-occurrence edge de-duplication, Bresenham
//   k_sky_setup + k_sky_fill   Framebuffer::render_skybox step 1 (render.rs:81-139, :242-299)
//   k_stars_claim + k_stars_write   render_skybox step 2: the star diamonds (render.rs:149-235)
//   k_transform    the transform alone (stage-output test hook); k_fb_clear, k_tex_expand, k_tex_mask, k_tex8_*: utilities
//
// Pixel-order semantics.  The reference draws surfaces one after the other, so a
// pixel's final value is a fold over the surfaces covering it, in draw order; pixels are independent.
// Every fill kernel therefore gives each pixel to one thread that keeps colour + depth in registers.
//   * Pass 1 never blends (has_transparency == false  =>  blend Opaque, editor_alpha 255), so its
//     fold has a closed form that needs NO sort:
//       painter's mode : the winner is the writing fragment drawn last = max (sort key, face index)
//       z-buffer mode  : the winner is the writing fragment with min (z, face index); `z < zbuf` is
//                        strict, so among equal depths the first-drawn (lowest index) wins, and the
//                        framebuffer's incoming depth wins every tie.
//     k_fill_opaque walks a tile's surfaces in whatever order the binning produced and keeps the
//     winner; the result is the reference's, bit for bit, and deterministic.
//   * Pass 2 blends against the running colour, so it is replayed in exact draw order: every surface carries a
//     unique 64-bit draw-order key (pass, depth key, face index), sorted per tile inside k_fill_ordered.
#include <algorithm>
#include <cstdlib>
#include <tuple>
#include <type_traits>
#include <utility>

#include "b32_device.cuh"

#include "b32_launch.h"

namespace b32 {

// Programmatic dependent launch: k_setup -> k_bin_opaque -> k_fill_opaque are chained on one stream; a
// dependent kernel is launched while its predecessor still runs, does the part of its prologue that reads
// nothing the predecessor writes, then waits here until the predecessor has completed and flushed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- HostStatus (b32_device.cuh): what a blocking call's kernels tell the host without a copy ----
__device__ __forceinline__ unsigned long long gtime64() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void host_stamp_start(const CallParams& p, int k) {
    if (p.host && blockIdx.x == 0 && threadIdx.x == 0) p.host->stamp[k].t0 = gtime64();
}
// Called by every thread of the CTA on each of the kernel's ways out.  The last CTA to arrive publishes.
__device__ __forceinline__ void host_signal_done(const CallParams& p, CallState* st, int k) {
    if (!p.host) return;
    __syncthreads();                                       // the CTA's writes (counters, flags) happen before thread 0's fence,
    if (threadIdx.x != 0) return;                          // which is cumulative: one fence orders them before the arrival
    if (blockIdx.x == 0) __threadfence_system(); else __threadfence();     // (block 0: t0 reaches the host before seq can)
    if (atomicAdd(&st->done[k], 1u) != gridDim.x - 1) return;
    __threadfence();
    HostStatus* h = p.host;
    if (k == HS_SETUP) {
        const volatile uint32_t* src = reinterpret_cast<const volatile uint32_t*>(st);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&h->state);
        #pragma unroll
        for (int i = 0; i < (int)(sizeof(CallState) / 4); ++i) dst[i] = src[i];
    }
    h->stamp[k].t1 = gtime64();
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(&h->stamp[k].seq) = p.host_seq;
}

// =================================================================================================
// vertex transform + snap (render.rs:2321-2360)
// =================================================================================================
__device__ __forceinline__ TVert transform_vertex(float px, float py, float pz, const CallParams& p, float* cam_xy) {
    // rel_pos = v.pos - camera.position; cam_pos = perspective_transform(...)   (math.rs:103-109)
    float rx = px - p.cam_pos[0], ry = py - p.cam_pos[1], rz = pz - p.cam_pos[2];
    float cx = rx * p.bx[0] + ry * p.bx[1] + rz * p.bx[2];
    float cy = rx * p.by[0] + ry * p.by[1] + rz * p.by[2];
    float cz = rx * p.bz[0] + ry * p.bz[1] + rz * p.bz[2];
    if (cam_xy) { cam_xy[0] = cx; cam_xy[1] = cy; }
    float sx, sy, sz;
    if (p.ortho) {                                       // math.rs:140-148
        sx = (cx - p.ortho_cx) * p.ortho_zoom + ((float)p.width / 2.0f);
        sy = -(cy - p.ortho_cy) * p.ortho_zoom + ((float)p.height / 2.0f);
        sz = cz;
    } else if (p.use_fixed_point) {                      // fixed.rs:362-441
        const int32_t distance = 5 * 4096, scale = 4 * 4096;     // Fixed32::from_f32(5.0), (4.0): fixed.rs:396-397
        int32_t wx = fx_from_f32(px), wy = fx_from_f32(py), wz = fx_from_f32(pz);
        int32_t ex = fx_sub(wx, p.fcam_pos[0]), ey = fx_sub(wy, p.fcam_pos[1]), ez = fx_sub(wz, p.fcam_pos[2]);
        int32_t fcx = fx_add(fx_add(fx_mul(ex, p.fbx[0]), fx_mul(ey, p.fbx[1])), fx_mul(ez, p.fbx[2]));
        int32_t fcy = fx_add(fx_add(fx_mul(ex, p.fby[0]), fx_mul(ey, p.fby[1])), fx_mul(ez, p.fby[2]));
        int32_t fcz = fx_add(fx_add(fx_mul(ex, p.fbz[0]), fx_mul(ey, p.fbz[1])), fx_mul(ez, p.fbz[2]));
        int32_t denom = fx_add(fcz, distance);
        int32_t adenom = denom < 0 ? (int32_t)(0u - (uint32_t)denom) : denom;   // release-mode i32::abs
        int32_t isx, isy;
        if (adenom < 256) {                              // fixed.rs:406-408
            isx = p.half_w >> 12; isy = p.half_h >> 12;
        } else {
            uint32_t nr2, shift;
            unr_recip(denom, &nr2, &shift);              // one reciprocal serves x and y
            int32_t proj_x = unr_apply(fx_mul(fcx, scale), denom, nr2, shift);
            int32_t proj_y = unr_apply(fx_mul(fcy, scale), denom, nr2, shift);
            isx = fx_add(fx_mul(proj_x, p.viewport_scale), p.half_w) >> 12;
            isy = fx_add(fx_mul(proj_y, p.viewport_scale), p.half_h) >> 12;
        }
        sx = (float)isx; sy = (float)isy;
        sz = cz + 5.0f;                                  // render.rs:2342-2345
    } else {                                             // math.rs:117-136
        const float us = 4.0f;
        float vs = ((float)min(p.width, p.height) / 2.0f) * 0.75f;
        float denom = cz + 5.0f;
        if (fabsf(denom) < 0.001f) {
            sx = (float)p.width / 2.0f; sy = (float)p.height / 2.0f; sz = cz;
        } else {
            sx = (cx * us) / denom * vs + ((float)p.width / 2.0f);
            sy = (cy * us) / denom * vs + ((float)p.height / 2.0f);
            sz = denom;
        }
    }
    return make_float4(sx, sy, sz, cz);
}

__global__ void __launch_bounds__(256)
k_transform(const b32_vertex* __restrict__ verts, TVert* __restrict__ out, float* __restrict__ dbg_cam, CallParams p) {
    pdl_launch_dependents();           // split front end (B32_SPLIT_TRANSFORM): k_setup is scheduled behind this grid and waits for its completion
    const uint32_t vw = p.vwords ? p.vwords : 9u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.nv; i += gridDim.x * blockDim.x) {
        const float* vp = reinterpret_cast<const float*>(verts) + (size_t)i * vw;
        float cxy[2];
        TVert t = transform_vertex(vp[0], vp[1], vp[2], p, cxy);
        out[i] = t;
        if (dbg_cam) { dbg_cam[i * 3] = cxy[0]; dbg_cam[i * 3 + 1] = cxy[1]; dbg_cam[i * 3 + 2] = t.w; }
    }
}

// ---- asynchronous copies ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// =================================================================================================
// k_setup
// =================================================================================================
struct V3 { float x, y, z; };
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 normalize3(V3 a) {                      // math.rs:39-49
    float l = sqrtf(dot3(a, a));
    if (l == 0.0f) return V3{0.0f, 0.0f, 0.0f};
    return V3{a.x / l, a.y / l, a.z / l};
}

// f32::acos as the reference's shipped wasm build computes it (compiler_builtins' libm `acosf`, a port of musl's
// e_acosf.c; render.rs:1047 calls it): plain f32 operators, one rounding each (the library is compiled with
// -fmad=false; sqrtf / division are the IEEE ones).
__device__ __forceinline__ float acosf_rpoly(float z) {
    float p = z * (0.16666586697101593f + z * (-0.04274342209100723f + z * -0.008656363002955914f));
    float q = z * -0.7066296339035034f + 1.0f;
    return p / q;
}
__device__ __forceinline__ float ref_acosf(float x) {
    const float pio2_hi = 1.570796251296997f, pio2_lo = 7.549789415861596e-08f;
    uint32_t hx = __float_as_uint(x), ix = hx & 0x7fffffffu;
    if (ix >= 0x3f800000u) {
        if (ix == 0x3f800000u) return (hx >> 31) ? 3.141592502593994f : 0.0f;
        return 0.0f / (x - x);
    }
    if (ix < 0x3f000000u) {
        if (ix <= 0x32800000u) return pio2_hi;
        return pio2_hi - (x - (pio2_lo - x * acosf_rpoly(x * x)));
    }
    if (hx >> 31) {
        float z = (1.0f + x) * 0.5f;
        float s = sqrtf(z);
        float w = acosf_rpoly(z) * s - pio2_lo;
        float t = pio2_hi - (s + w);
        return t + t;
    }
    float z = (1.0f - x) * 0.5f;
    float s = sqrtf(z);
    float df = __uint_as_float(__float_as_uint(s) & 0xfffff000u);
    float c = (z - df * df) / (s + df);
    float w = acosf_rpoly(z) * s + c;
    float t = df + w;
    return t + t;
}

// render.rs:1013-1071 (Directional, Point, Spot).  SPOT = false compiles the kernel the usual scenes run (the app only
// constructs point lights, scene.rs:62): the acos path costs registers in k_setup, so calls with an enabled Spot light
// run a second instantiation.
template <bool SPOT>
__device__ void shade_multi_light_color(V3 n, V3 wp, const LightDev* __restrict__ lights, uint32_t nl, float ambient, float* out) {
    float tr = ambient, tg = ambient, tb = ambient;
    for (uint32_t i = 0; i < nl; ++i) {
        LightDev L = lights[i];
        if (!L.enabled) continue;
        float contribution;
        if (L.type == B32_LIGHT_DIRECTIONAL) {
            V3 neg{L.dx * -1.0f, L.dy * -1.0f, L.dz * -1.0f};
            float ndl = fmaxf(dot3(n, neg), 0.0f);
            contribution = ndl * L.intensity;
        } else {
            V3 to_light{L.px - wp.x, L.py - wp.y, L.pz - wp.z};
            float dist = sqrtf(dot3(to_light, to_light));
            if (dist > L.radius || dist < 0.001f) {
                contribution = 0.0f;
            } else {
                V3 tl = normalize3(to_light);
                float edge = 1.0f;
                bool lit = true;
                if (SPOT && L.type == B32_LIGHT_SPOT) {                           // render.rs:1045-1053
                    V3 neg{tl.x * -1.0f, tl.y * -1.0f, tl.z * -1.0f};
                    float spot_angle = ref_acosf(dot3(neg, V3{L.dx, L.dy, L.dz}));
                    lit = !(spot_angle > L.angle);
                    edge = 1.0f - (spot_angle / L.angle);
                }
                if (!lit) {
                    contribution = 0.0f;
                } else {
                    float att = 1.0f - (dist / L.radius);
                    float ndl = fmaxf(dot3(n, tl), 0.0f);
                    contribution = ndl * L.intensity * att * att;
                    if (SPOT && L.type == B32_LIGHT_SPOT) contribution = contribution * edge;
                }
            }
        }
        tr += contribution * L.cr; tg += contribution * L.cg; tb += contribution * L.cb;
    }
    out[0] = fminf(tr, 1.0f); out[1] = fminf(tg, 1.0f); out[2] = fminf(tb, 1.0f);
}

// render.rs:2266-2293
__device__ __forceinline__ uint32_t fog_color(uint32_t c, float z, const CallParams& p) {
    float f;
    if (z <= p.fog_start) f = 0.0f;
    else if (p.fog_falloff <= 0.0f) f = 1.0f;
    else f = fminf((z - p.fog_start) / p.fog_falloff, 1.0f);
    if (f <= 0.0f) return c;
    if (f >= 1.0f) return (uint32_t)p.fog_r | ((uint32_t)p.fog_g << 8) | ((uint32_t)p.fog_b << 16) | ((uint32_t)p.fog_blend << 24);
    float inv = 1.0f - f;
    uint32_t r = f2u8((float)(c & 0xFF) * inv + (float)p.fog_r * f);
    uint32_t g = f2u8((float)((c >> 8) & 0xFF) * inv + (float)p.fog_g * f);
    uint32_t b = f2u8((float)((c >> 16) & 0xFF) * inv + (float)p.fog_b * f);
    return r | (g << 8) | (b << 16);             // Color::new => blend Opaque (0)
}

__device__ __forceinline__ bool is_integral(float x) { return truncf(x) == x; }

// One thread per face.  `tv` (pre-transformed vertices) may be NULL: then the three vertices are
// transformed here (no intermediate vertex buffer: each vertex record is read once per use).
// Pass-1 surfaces also emit a 16-byte BinHead (heads[face]); k_bin_opaque scatters those to tiles.
#ifndef B32_SETUP_THREADS
#define B32_SETUP_THREADS 128
#endif
constexpr int SETUP_THREADS = B32_SETUP_THREADS;

template <bool STAGED, bool SPOT>
__device__ __forceinline__ void setup_face(uint32_t fi, const uint4& fc, const float* __restrict__ gverts, const float* __restrict__ s_vert, uint32_t lo,
                                           const TVert* __restrict__ tv, const TexDev* __restrict__ tex,
                                           const LightDev* __restrict__ lights,
                                           SurfRec* __restrict__ recs, uint64_t* __restrict__ keys,
                                           CallState* __restrict__ st, const CallParams& p,
                                           uint32_t& n_op, uint32_t& n_tr, BinHead& head, bool& binned,
                                           bool& marked, uint32_t& mbx, uint32_t& mby,
                                           WireTri* __restrict__ wire) {
    binned = false;
    if (wire) wire[fi].kind = 0;
    uint32_t cls = 2;                 // 0 opaque pass, 1 transparent pass, 2 not drawn
    uint32_t dkey = 0;
    do {
        if (fc.x >= p.nv || fc.y >= p.nv || fc.z >= p.nv) { st->oob = 1; break; }
        uint32_t tex_id = fc.w & 0xFFFFu, face_blend = (fc.w >> 16) & 7u, editor_alpha = fc.w >> 24;
        if (face_blend > B32_BLEND_ERASE) { st->oob = 2; break; }             // not a BlendMode (types.rs): the call draws nothing, B32_ERR_INVALID
        bool black_tr = (fc.w >> 19) & 1u;
        bool textured = tex_id != B32_FACE_TEX_NONE && tex_id < p.ntex;
        uint32_t tex_blend = 0;
        if (textured) tex_blend = tex[tex_id].blend;

        // vertex records of p.vwords words (9 = b32_vertex, 6 = b32_vertex_nn without the normal); staged groups read them
        // from the shared-memory window [lo, lo + SETUP_WIN)
        const uint32_t vw = p.vwords, cw = vw - 1;                             // the colour is the last word
        const float* v0 = STAGED ? s_vert + (fc.x - lo) * vw : gverts + (size_t)fc.x * vw;
        const float* v1 = STAGED ? s_vert + (fc.y - lo) * vw : gverts + (size_t)fc.y * vw;
        const float* v2 = STAGED ? s_vert + (fc.z - lo) * vw : gverts + (size_t)fc.z * vw;
        TVert t1, t2, t3;
        if (tv) { t1 = tv[fc.x]; t2 = tv[fc.y]; t3 = tv[fc.z]; }
        else {
            t1 = transform_vertex(v0[0], v0[1], v0[2], p, nullptr);
            t2 = transform_vertex(v1[0], v1[1], v1[2], p, nullptr);
            t3 = transform_vertex(v2[0], v2[1], v2[2], p, nullptr);
        }
        if (!p.ortho) {                                                       // :2380-2385
            if (t1.w <= NEAR_PLANE || t2.w <= NEAR_PLANE || t3.w <= NEAR_PLANE) break;
        }
        float signed_area = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);   // :2393
        bool backface = signed_area <= 0.0f;
        // render_mesh_15: has_transparency (:2403-2415) = drawn in pass 2.  render_mesh (RGB888) has ONE list and a blend
        // tag per texel, so here the flag means "this surface may read the framebuffer" (its texture holds blended
        // texels — TexDev.blend of the RGB888 table — or editor_alpha < 255): any such surface sends the whole call
        // through the strict draw-order replay.
        bool transparent;
        if (p.rgb888) transparent = (textured && tex_blend != 0) || editor_alpha < 255;
        else if (textured && tex_blend != B32_BLEND_OPAQUE) transparent = true;
        else if (face_blend != B32_BLEND_OPAQUE) transparent = true;
        else transparent = editor_alpha < 255;
        if (p.fog_enabled && t1.w > p.fog_cull && t2.w > p.fog_cull && t3.w > p.fog_cull) break;   // :2421-2424
        if (wire) {       // wireframe phase inputs: back faces unless x-ray (:2446-2450), front faces in overlay mode (:2509-2511)
            uint32_t kind = backface ? ((p.wire_back && !p.xray_mode) ? 1u : 0u) : (p.wire_front ? 2u : 0u);
            if (kind) wire[fi] = WireTri{{t1.x, t2.x, t3.x}, {t1.y, t2.y, t3.y}, {t1.z, t2.z, t3.z}, kind};
        }
        if (backface && !(!p.backface_cull || p.xray_mode)) break;            // :2445-2453

        // vertex attributes (36-byte records: pos 0, uv 12, normal 20, rgba 32); v2/v3 swap :2455-2457
        TVert s1 = t1, s2 = backface ? t3 : t2, s3 = backface ? t2 : t3;
        const float* va = v0;
        const float* vb = backface ? v2 : v1;
        const float* vc = backface ? v1 : v2;
        uint32_t c1 = reinterpret_cast<const uint32_t*>(va)[cw], c2 = reinterpret_cast<const uint32_t*>(vb)[cw], c3 = reinterpret_cast<const uint32_t*>(vc)[cw];
        if (p.fog_enabled) {                                                  // :2427-2436 (cam z of the same vertex)
            c1 = fog_color(c1, s1.w, p); c2 = fog_color(c2, s2.w, p); c3 = fog_color(c3, s3.w, p);
        }

        SurfRec r;
        uint32_t blend_mode = textured ? tex_blend : face_blend;               // :1450-1452
        if (p.rgb888) { blend_mode = 0; black_tr = textured; }                 // RGB888: blend tags are per texel; Erase texels skip (:1350)
        uint32_t flags = blend_mode | (black_tr ? SF_BLACK_TR : 0) | (textured ? SF_TEXTURED : 0) |
                         (transparent ? SF_TRANSPARENT : 0) | (editor_alpha << 8) | ((textured ? tex_id : 0xFFFFu) << 16);
        bool needs_dither = p.dithering && (p.shading == B32_SHADE_GOURAUD || textured || c1 != c2 || c2 != c3);   // :1487-1492
        if (needs_dither) flags |= SF_DITHER;

        // bounding box, render.rs:1455-1463
        uint32_t min_x = f2u32sat(fmaxf(fminf(fminf(s1.x, s2.x), s3.x), 0.0f));
        uint32_t max_x = f2u32sat(fminf(fmaxf(fmaxf(s1.x, s2.x), s3.x) + 1.0f, (float)p.width));
        uint32_t min_y = f2u32sat(fmaxf(fminf(fminf(s1.y, s2.y), s3.y), 0.0f));
        uint32_t max_y = f2u32sat(fminf(fmaxf(fmaxf(s1.y, s2.y), s3.y) + 1.0f, (float)p.height));
        bool empty = min_x >= max_x || min_y >= max_y;
        float area = (s2.y - s3.y) * (s1.x - s3.x) + (s3.x - s2.x) * (s1.y - s3.y);     // :1500
        if (fabsf(area) < 0.00001f) empty = true;                                      // :1501-1503
        if (empty) { min_x = max_x = min_y = max_y = 0; }
        r.inv_area = 1.0f / area;
        r.a0 = s2.y - s3.y; r.b0 = s3.x - s2.x; r.a1 = s3.y - s1.y; r.b1 = s1.x - s3.x;   // :1507-1510
        float start_x = (float)min_x, start_y = (float)min_y;
        r.w0s = r.a0 * (start_x - s3.x) + r.b0 * (start_y - s3.y);                      // :1517-1518
        r.w1s = r.a1 * (start_x - s3.x) + r.b1 * (start_y - s3.y);
        r.bbox_x = min_x | (max_x << 16);
        r.bbox_y = min_y | (max_y << 16);
        // incremental stepping == closed form when everything is an integer below 2^23
        {
            float nx = (float)(max_x - min_x), ny = (float)(max_y - min_y);
            float m0 = fabsf(r.w0s) + ny * fabsf(r.b0) + nx * fabsf(r.a0);
            float m1 = fabsf(r.w1s) + ny * fabsf(r.b1) + nx * fabsf(r.a1);
            bool ints = is_integral(r.a0) && is_integral(r.b0) && is_integral(r.a1) && is_integral(r.b1) &&
                        is_integral(r.w0s) && is_integral(r.w1s);
            if (ints && m0 < 8388608.0f && m1 < 8388608.0f) flags |= SF_FAST_EDGE;
            else if (!empty && !fill_uses_edge_prefix(p) && (max_x - min_x) * (max_y - min_y) * 64u >= p.width * p.height) {
                // a fixed-point call's large surface with stepped edge values (far off-screen vertices): counted (one atomic per
                // warp that gets here); the fill kernel tells the host, which picks the fill instantiation of the NEXT calls by it
                const uint32_t am = __activemask();
                if ((threadIdx.x & 31u) == (uint32_t)(__ffs(am) - 1)) atomicAdd(&st->n_big_stepped, __popc(am));
            }
        }
        r.iz1 = 1.0f / s1.z; r.iz2 = 1.0f / s2.z; r.iz3 = 1.0f / s3.z;                  // :1546-1548
        r.u1 = va[3]; r.v1 = va[4]; r.u2 = vb[3]; r.v2 = vb[4]; r.u3 = vc[3]; r.v3 = vc[4];
        r.vc1 = c1 & 0xFFFFFF; r.vc2 = c2 & 0xFFFFFF; r.vc3 = c3 & 0xFFFFFF;
        for (int k = 0; k < 9; ++k) r.sh[k] = 1.0f;
        if (!empty && p.shading != B32_SHADE_NONE) {
            V3 w1{va[0], va[1], va[2]}, w2{vb[0], vb[1], vb[2]}, w3{vc[0], vc[1], vc[2]};
            V3 n1{va[5], va[6], va[7]}, n2{vb[5], vb[6], vb[7]}, n3{vc[5], vc[6], vc[7]};
            if (backface) {                                                             // wn.scale(-1.0) :2464-2466
                n1 = V3{n1.x * -1.0f, n1.y * -1.0f, n1.z * -1.0f}; n2 = V3{n2.x * -1.0f, n2.y * -1.0f, n2.z * -1.0f};
                n3 = V3{n3.x * -1.0f, n3.y * -1.0f, n3.z * -1.0f};
            }
            if (p.shading == B32_SHADE_FLAT) {                                          // :1466-1472
                const float third = 1.0f / 3.0f;
                V3 c{((w1.x + w2.x) + w3.x) * third, ((w1.y + w2.y) + w3.y) * third, ((w1.z + w2.z) + w3.z) * third};
                V3 n = normalize3(V3{((n1.x + n2.x) + n3.x) * third, ((n1.y + n2.y) + n3.y) * third, ((n1.z + n2.z) + n3.z) * third});
                shade_multi_light_color<SPOT>(n, c, lights, p.n_lights, p.ambient, r.sh);
            } else {                                                                    // :1475-1483
                shade_multi_light_color<SPOT>(n1, w1, lights, p.n_lights, p.ambient, r.sh);
                shade_multi_light_color<SPOT>(n2, w2, lights, p.n_lights, p.ambient, r.sh + 3);
                shade_multi_light_color<SPOT>(n3, w3, lights, p.n_lights, p.ambient, r.sh + 6);
            }
        }
        r.flags = flags;
        r._pad = 0;
        recs[fi] = r;

        cls = transparent ? 1u : 0u;
        float center_z = (s1.z + s2.z + s3.z) / 3.0f;                                   // :2529
        bool sorted = (transparent && !p.rgb888) || !p.use_zbuffer;                     // :2527, :2536; RGB888: :2155
        if (sorted) {
            if (center_z != center_z) { if (transparent && !p.rgb888) st->nan_transp = 1; else st->nan_opaque = 1; }
            dkey = depth_key_desc(center_z);
        }
        if (transparent) ++n_tr; else ++n_op;

        // pass-1 surfaces go into their screen tiles' bins (any order; see file header)
        if (!transparent && !(p.xray_mode && !p.rgb888) && !empty) {
            uint32_t hkey = dkey;
            if (p.use_zbuffer) {
                // front-to-back walk key: a lower bound of every depth this surface can produce.
                // 1/z = sum(bc_i / z_i) with bc_i >= -1e-4 and sum(bc) = 1.  If all z_i > 0 and
                // zmax <= 1000 * zmin then 0 < 1/z <= (1 + 2e-4) / zmin, so z >= 0.9997 * zmin; the bound
                // 0.999 * zmin also absorbs rounding.  Otherwise (ortho depths <= 0, NaN, extreme depth
                // ratios where 1/z can change sign) no bound is claimed: key 0xFFFFFFFF = "never cull".
                float zmin = fminf(fminf(s1.z, s2.z), s3.z), zmax = fmaxf(fmaxf(s1.z, s2.z), s3.z);
                bool ok = s1.z > 0.0f && s2.z > 0.0f && s3.z > 0.0f && zmax <= 1000.0f * zmin && zmax <= 3.0e38f;
                float lb = ok ? zmin * 0.999f : 0.0f;
                hkey = ~__float_as_uint(lb);           // lb >= 0: bits are monotone; inverted so "descending key" = front first
            }
            head = BinHead{r.bbox_x, r.bbox_y, hkey, fi};
            binned = true;
        }
        // every drawn, non-empty surface is marked in the tile masks: pass 1 reads the candidates whose head is set; pass 2
        // and x-ray mode (k_fill_ordered) those whose keys[] class says so, and rebuilds their draw-order key from
        // keys[fi] (pass, depth key) and this record's bbox.
        if (!empty) { marked = true; mbx = r.bbox_x; mby = r.bbox_y; }
    } while (0);
    keys[fi] = ((uint64_t)cls << 32) | dkey;
}

// Exact trivial reject of a surface against the pixel box [bx0,bx1) x [by0,by1) (it must overlap the surface's bbox).
// For SF_FAST_EDGE surfaces the edge values are exact integers, linear in (x, y), and bc = fl(w * inv_area) is monotone
// in w, so each barycentric's extreme over the box (clipped to the bbox) is at a corner; bc_z = fl(fl(1 - bc_x) - bc_y) is
// monotone non-increasing in both.  A surface whose upper bounds fail `>= -0.0001` (render.rs:1541) has no inside pixel
// in the box.  Surfaces without the flag (rounded stepping) are never rejected here.
template <typename Rec>
__device__ __forceinline__ bool surface_misses_box(const Rec& r, uint32_t bx0, uint32_t bx1, uint32_t by0, uint32_t by1) {
    if (!(r.flags & SF_FAST_EDGE)) return false;
    const uint32_t min_x = r.bbox_x & 0xFFFF, max_x = r.bbox_x >> 16, min_y = r.bbox_y & 0xFFFF, max_y = r.bbox_y >> 16;
    float dx0 = (float)(max(bx0, min_x) - min_x), dx1 = (float)(min(bx1, max_x) - 1 - min_x);
    float dy0 = (float)(max(by0, min_y) - min_y), dy1 = (float)(min(by1, max_y) - 1 - min_y);
    float r0 = r.w0s + dy0 * r.b0, r1 = r.w0s + dy1 * r.b0, q0 = r.w1s + dy0 * r.b1, q1 = r.w1s + dy1 * r.b1;
    float ax0 = dx0 * r.a0, ax1 = dx1 * r.a0, cx0 = dx0 * r.a1, cx1 = dx1 * r.a1;
    float x00 = (r0 + ax0) * r.inv_area, x01 = (r0 + ax1) * r.inv_area, x10 = (r1 + ax0) * r.inv_area, x11 = (r1 + ax1) * r.inv_area;
    float y00 = (q0 + cx0) * r.inv_area, y01 = (q0 + cx1) * r.inv_area, y10 = (q1 + cx0) * r.inv_area, y11 = (q1 + cx1) * r.inv_area;
    float xmax = fmaxf(fmaxf(x00, x01), fmaxf(x10, x11)), xmin = fminf(fminf(x00, x01), fminf(x10, x11));
    float ymax = fmaxf(fmaxf(y00, y01), fmaxf(y10, y11)), ymin = fminf(fminf(y00, y01), fminf(y10, y11));
    const float ERR = -0.0001f;
    return xmax < ERR || ymax < ERR || (1.0f - xmin - ymin) < ERR;
}

// The same question for a surface WITHOUT SF_FAST_EDGE (its edge values are the reference's chain of rounded additions,
// not a closed form): a conservative answer.  Let L(dx, dy) = w_start + dy * b + dx * a in real arithmetic.  Each of the
// n = dx + dy chain steps rounds once (relative error <= 2^-24 of a partial sum, and every partial sum is bounded by
// M = |w_start| + dy_max |b| + dx_max |a| up to (1 + 2^-24)^n), so |chain - L| <= n 2^-23 M; the float evaluation of L at a
// corner is within 4 * 2^-24 M.  E = (n + 8) 2^-22 M covers both twice over.  L is linear, so its extremes over the box are
// at the corners; bc = fl(chain * inv_area) adds one more rounding, covered by the slack term.  The surface is rejected
// only when an upper bound of a barycentric over the whole box is below the inside threshold (render.rs:1541): every
// pixel of the box then fails the exact test, so dropping the surface cannot change the frame.  Non-finite values make
// every comparison false: no reject.  tests/test_edge_reject.py holds the numpy mirror, checked against brute-force chains.
template <typename Rec>
__device__ __forceinline__ bool stepped_surface_misses_box(const Rec& r, uint32_t bx0, uint32_t bx1, uint32_t by0, uint32_t by1) {
    const uint32_t min_x = r.bbox_x & 0xFFFF, max_x = r.bbox_x >> 16, min_y = r.bbox_y & 0xFFFF, max_y = r.bbox_y >> 16;
    const uint32_t ix0 = max(bx0, min_x) - min_x, ix1 = min(bx1, max_x) - 1 - min_x, iy0 = max(by0, min_y) - min_y, iy1 = min(by1, max_y) - 1 - min_y;
    const float dx0 = (float)ix0, dx1 = (float)ix1, dy0 = (float)iy0, dy1 = (float)iy1;
    const float r0 = r.w0s + dy0 * r.b0, r1 = r.w0s + dy1 * r.b0, q0 = r.w1s + dy0 * r.b1, q1 = r.w1s + dy1 * r.b1;
    const float ax0 = dx0 * r.a0, ax1 = dx1 * r.a0, cx0 = dx0 * r.a1, cx1 = dx1 * r.a1;
    const float x00 = (r0 + ax0) * r.inv_area, x01 = (r0 + ax1) * r.inv_area, x10 = (r1 + ax0) * r.inv_area, x11 = (r1 + ax1) * r.inv_area;
    const float y00 = (q0 + cx0) * r.inv_area, y01 = (q0 + cx1) * r.inv_area, y10 = (q1 + cx0) * r.inv_area, y11 = (q1 + cx1) * r.inv_area;
    const float xmax = fmaxf(fmaxf(x00, x01), fmaxf(x10, x11)), xmin = fminf(fminf(x00, x01), fminf(x10, x11));
    const float ymax = fmaxf(fmaxf(y00, y01), fmaxf(y10, y11)), ymin = fminf(fminf(y00, y01), fminf(y10, y11));
    const float steps = (float)(ix1 + iy1 + 8u) * 2.384185791015625e-07f;                                  // (n + 8) 2^-22
    const float ia = fabsf(r.inv_area);
    const float ex = steps * (fabsf(r.w0s) + dy1 * fabsf(r.b0) + dx1 * fabsf(r.a0)) * ia;
    const float ey = steps * (fabsf(r.w1s) + dy1 * fabsf(r.b1) + dx1 * fabsf(r.a1)) * ia;
    const float slack = 9.5367431640625e-07f * (1.0f + fabsf(xmax) + fabsf(xmin) + fabsf(ymax) + fabsf(ymin));   // 2^-20 (...)
    const float ERR = -0.0001f;
    return xmax + ex + slack < ERR || ymax + ey + slack < ERR || (1.0f - xmin - ymin) + ex + ey + 3.0f * slack < ERR;
}

// ---- tile masks (see b32_device.cuh: "Binning without bins") ----------------------------------------
__device__ __forceinline__ void bbox_mtiles(uint32_t bbox_x, uint32_t bbox_y, uint32_t shift, uint32_t& tx0, uint32_t& tx1, uint32_t& ty0, uint32_t& ty1) {
    uint32_t min_x = bbox_x & 0xFFFF, max_x = bbox_x >> 16, min_y = bbox_y & 0xFFFF, max_y = bbox_y >> 16;
    tx0 = min_x >> shift; tx1 = (max_x - 1) >> shift; ty0 = min_y >> shift; ty1 = (max_y - 1) >> shift;
}

// f(mask tile, owner thread) for every mask tile the bounding box of each lane's head touches.  Must be called by all
// 32 lanes of a warp (has = this lane holds a head).  A head that touches few tiles is walked by its own lane; one that
// touches many (a wall close to the camera covers hundreds of tiles) is walked by the whole warp, 32 tiles at a time,
// so no lane ever runs a long serial loop while the other 31 wait.
constexpr uint32_t COOP_TILES = 16;
template <typename F>
__device__ __forceinline__ void for_each_mtile(uint32_t bbox_x, uint32_t bbox_y, bool has, uint32_t pix_shift, uint32_t mtiles_x, F f) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t tx0 = 0, tx1 = 0, ty0 = 0, ty1 = 0, nt = 0;
    if (has) { bbox_mtiles(bbox_x, bbox_y, pix_shift, tx0, tx1, ty0, ty1); nt = (tx1 - tx0 + 1) * (ty1 - ty0 + 1); }
    const bool big = nt > COOP_TILES;
    if (has && !big)
        for (uint32_t ty = ty0; ty <= ty1; ++ty)
            for (uint32_t tx = tx0; tx <= tx1; ++tx) f(ty * mtiles_x + tx, threadIdx.x);
    uint32_t bigmask = __ballot_sync(0xFFFFFFFFu, big);
    while (bigmask) {
        const int src = __ffs(bigmask) - 1;
        bigmask &= bigmask - 1;
        const uint32_t bx0 = __shfl_sync(0xFFFFFFFFu, tx0, src), by0 = __shfl_sync(0xFFFFFFFFu, ty0, src);
        const uint32_t bw = __shfl_sync(0xFFFFFFFFu, tx1, src) - bx0 + 1, total = __shfl_sync(0xFFFFFFFFu, nt, src);
        for (uint32_t i = lane; i < total; i += 32) f((by0 + i / bw) * mtiles_x + bx0 + i % bw, (threadIdx.x & ~31u) + (uint32_t)src);
    }
}

// One group's tile masks: zero, mark (every thread of the CTA calls mark with its own face's bounding box), flush to
// masks[mask tile][group].  blockDim.x == SETUP_GROUP.
__device__ __forceinline__ void masks_zero(uint4* s_mask, uint32_t n_mtiles) {
    for (uint32_t i = threadIdx.x; i < n_mtiles; i += blockDim.x) s_mask[i] = make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void masks_mark(uint4* s_mask, uint32_t bbox_x, uint32_t bbox_y, bool has, const CallParams& p) {
    uint32_t* w = reinterpret_cast<uint32_t*>(s_mask);
    for_each_mtile(bbox_x, bbox_y, has, 4u + p.mshift, p.mtiles_x, [&](uint32_t t, uint32_t owner) { atomicOr(&w[t * 4 + (owner >> 5)], 1u << (owner & 31)); });
}
// ord = which faces of the group are replayed by the ordered pass (pass 2 / x-ray / RGB888): their number per mask tile is
// accumulated in ocount[] and its maximum in st->obin_max, so that k_fill_ordered knows BEFORE any tile draws whether every
// tile's draw-order entries fit (shared memory, or its slice of the scratch).  Ordinary frames have none: no atomics.
__device__ __forceinline__ void masks_flush(const uint4* s_mask, uint4* __restrict__ masks, uint32_t n_mtiles, uint32_t group, const CallParams& p,
                                            uint4 ord, uint32_t* __restrict__ ocount, CallState* __restrict__ st) {
    const bool any_ord = (ord.x | ord.y | ord.z | ord.w) != 0;
    for (uint32_t t = threadIdx.x; t < n_mtiles; t += blockDim.x) {
        const uint4 m = s_mask[t];
        masks[(size_t)t * p.n_groups + group] = m;
        if (any_ord) {
            const uint32_t c = __popc(m.x & ord.x) + __popc(m.y & ord.y) + __popc(m.z & ord.z) + __popc(m.w & ord.w);
            if (c) atomicMax(&st->obin_max, atomicAdd(&ocount[t], c) + c);
        }
    }
}

// ---- a fill CTA's surface list out of its mask row ----------------------------------------------------
// Thread t owns the groups t, t + THREADS, t + 2 THREADS, ... of the row (a warp reads 512 contiguous bytes per load),
// and its candidates take consecutive list positions: the list is a fixed enumeration of the set bits, not in face order
// (no user needs one: pass 1 is order-free, the ordered pass sorts by key, the skybox compares face indices).
// cand_scan counts the set bits (block-wide exclusive scan); cand_expand writes the face indices of the candidates
// with list position in [w0, w0 + wn) to out[position - w0].  Both contain barriers: every thread of the CTA must call them.
struct CandScan { uint32_t n_groups, stride, base, total; };
__device__ __forceinline__ uint32_t popc128(const uint4& m) { return __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w); }

template <int THREADS>
__device__ __forceinline__ CandScan cand_scan(const uint4* __restrict__ mrow, uint32_t n_groups, uint32_t* s_wsum /* [THREADS / 32] */) {
    CandScan cs;
    cs.n_groups = n_groups; cs.stride = THREADS;
    uint32_t cnt = 0;
    #pragma unroll 4
    for (uint32_t g = threadIdx.x; g < n_groups; g += THREADS) cnt += popc128(mrow[g]);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t xs = cnt;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, xs, o); if (lane >= (uint32_t)o) xs += t; }
    __syncthreads();                                   // s_wsum may still be read by an earlier use
    if (lane == 31) s_wsum[warp] = xs;
    __syncthreads();
    uint32_t pre = 0, total = 0;
    for (uint32_t w = 0; w < (uint32_t)(THREADS / 32); ++w) { uint32_t v = s_wsum[w]; if (w < warp) pre += v; total += v; }
    cs.base = pre + xs - cnt;
    cs.total = total;
    return cs;
}

__device__ __forceinline__ void cand_expand(const uint4* __restrict__ mrow, const CandScan& cs, uint32_t w0, uint32_t wn, uint32_t* out) {
    uint32_t idx = cs.base;
    for (uint32_t g = threadIdx.x; g < cs.n_groups && idx < w0 + wn; g += 4 * cs.stride) {      // mask loads four at a time: one L2 latency per batch
        uint4 m[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = g + j * cs.stride < cs.n_groups ? mrow[g + j * cs.stride] : make_uint4(0, 0, 0, 0);
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = popc128(m[j]);
            if (c == 0) continue;
            if (idx + c <= w0 || idx >= w0 + wn) { idx += c; continue; }
            const uint32_t words[4] = {m[j].x, m[j].y, m[j].z, m[j].w};
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t bits = words[q];
                while (bits) {
                    const uint32_t b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (idx >= w0 && idx < w0 + wn) out[idx - w0] = (g + j * cs.stride) * SETUP_GROUP + q * 32 + b;
                    ++idx;
                }
            }
        }
    }
    __syncthreads();
}

#ifndef B32_SETUP_MINB
#define B32_SETUP_MINB 6          // <= 80 registers: 6 CTAs per SM = 888 slots, so the 782 CTAs of a 100k-face mesh are one wave
#endif
constexpr int SETUP_WIN = 3 * SETUP_THREADS;            // vertices staged per group when its indices fit a window this wide
constexpr size_t SETUP_STAGE_BYTES = (size_t)SETUP_WIN * sizeof(b32_vertex) + 32;
static_assert(SETUP_THREADS == SETUP_GROUP && SETUP_GROUP == 128, "one face per thread and group; a mask entry is one uint4");

// Dynamic shared memory: [n_mtiles] uint4 tile masks, then the staged vertex window.
// Vertex staging: when the three indices of all faces of the group lie in a window of SETUP_WIN vertices (always for an
// unindexed triangle soup, usually for level geometry), the window is copied with coalesced 16-byte cp.async pieces and
// the threads read their 27 floats from shared memory at a 27-word stride (conflict-free); otherwise they gather from
// global memory.  Every vertex buffer of the library is padded by 16 bytes so that the last piece may overrun the window.
template <bool STAGED, bool SPOT>
__device__ __forceinline__ void setup_group(uint32_t group, const b32_vertex* __restrict__ verts, const uint4& fc, const float* __restrict__ s_vert,
                                            uint32_t lo, const TVert* __restrict__ tv, const TexDev* __restrict__ tex, const LightDev* __restrict__ lights,
                                            SurfRec* __restrict__ recs, uint64_t* __restrict__ keys, BinHead* __restrict__ heads, uint4* s_mask,
                                            WireTri* __restrict__ wire, CallState* __restrict__ st, const CallParams& p, uint32_t& n_op, uint32_t& n_tr,
                                            uint32_t* s_ord) {
    const uint32_t fi = group * SETUP_GROUP + threadIdx.x;
    BinHead head{0, 0, 0, fi};                               // bbox 0 = not drawn in pass 1
    bool binned = false, marked = false;
    uint32_t mbx = 0, mby = 0;
    if (fi < p.nf) {
        setup_face<STAGED, SPOT>(fi, fc, reinterpret_cast<const float*>(verts), s_vert, lo, tv, tex, lights, recs, keys, st, p, n_op, n_tr, head, binned, marked, mbx, mby, wire);
        heads[fi] = head;
    }
    masks_mark(s_mask, mbx, mby, marked, p);
    // marked and not a pass-1 entry = replayed by the ordered pass; RGB888: any marked surface may be (one list)
    const uint32_t ob = __ballot_sync(0xFFFFFFFFu, marked && (!binned || p.rgb888));
    if ((threadIdx.x & 31) == 0) s_ord[threadIdx.x >> 5] = ob;
}

template <bool SPOT>
__global__ void __launch_bounds__(SETUP_THREADS, B32_SETUP_MINB)
k_setup(const b32_vertex* __restrict__ verts, const b32_face* __restrict__ faces, const TVert* __restrict__ tv,
        const TexDev* __restrict__ tex, const LightDev* __restrict__ lights,
        SurfRec* __restrict__ recs, uint64_t* __restrict__ keys,
        BinHead* __restrict__ heads, uint4* __restrict__ masks, uint32_t* __restrict__ ocount, WireTri* __restrict__ wire, CallState* __restrict__ st,
        uint32_t* __restrict__ zero_next, uint32_t zero_words,
        uint32_t* __restrict__ clear_rgba, float* __restrict__ clear_z, uint32_t clear_n, uint32_t clear_color, CallParams p) {
    extern __shared__ __align__(16) uint8_t su_smem[];
    __shared__ uint32_t s_cnt[2];
    __shared__ uint32_t s_lohi[2];
    __shared__ uint32_t s_ord[SETUP_THREADS / 32];
    const uint32_t n_mtiles = p.mtiles_x * p.mtiles_y;
    uint4* s_mask = reinterpret_cast<uint4*>(su_smem);
    uint8_t* s_stage = su_smem + (size_t)n_mtiles * sizeof(uint4);
    pdl_launch_dependents();           // the fill may be scheduled as SM resources free up; it waits for this grid's completion
    host_stamp_start(p, HS_SETUP);
    // Framebuffer::clear of the same frame (render.rs:36-45), folded in: nothing of this kernel reads the framebuffer, and
    // the previous frame's kernels have completed (this kernel is an ordinary, fully ordered launch)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < clear_n; i += gridDim.x * blockDim.x) { clear_rgba[i] = clear_color; clear_z[i] = 3.40282347e+38f; }
    // the call after this one finds its CallState zeroed (two sets, used alternately)
    if (blockIdx.x == 0) for (uint32_t i = threadIdx.x; i < zero_words; i += blockDim.x) zero_next[i] = 0;
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    uint32_t n_op = 0, n_tr = 0;
    const uint32_t lane = threadIdx.x & 31;
    if (tv) pdl_wait();                // split front end: k_transform (an ordinary, fully ordered launch) has completed
    for (uint32_t group = blockIdx.x; group < p.n_groups; group += gridDim.x) {
        const uint32_t fi = group * SETUP_GROUP + threadIdx.x;
        masks_zero(s_mask, n_mtiles);
        if (threadIdx.x == 0) { s_lohi[0] = 0xFFFFFFFFu; s_lohi[1] = 0; }
        uint4 fc = make_uint4(0, 0, 0, 0);
        uint32_t lo = 0xFFFFFFFFu, hi = 0;
        if (fi < p.nf) {
            // implicit faces: an unindexed triangle soup sends only the flags word; face i = vertices 3i, 3i+1, 3i+2
            if (p.faces_implicit) fc = make_uint4(3u * fi, 3u * fi + 1u, 3u * fi + 2u, p.faces_implicit == 2 ? p.uniform_flags : reinterpret_cast<const uint32_t*>(faces)[fi]);
            else fc = *reinterpret_cast<const uint4*>(faces + fi);
            lo = min(fc.x, min(fc.y, fc.z)); hi = max(fc.x, max(fc.y, fc.z));
        }
        lo = __reduce_min_sync(0xFFFFFFFFu, lo); hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        __syncthreads();                                   // masks zeroed, s_lohi initialised (and the previous group's flush is done)
        if (lane == 0) { atomicMin(&s_lohi[0], lo); atomicMax(&s_lohi[1], hi); }
        __syncthreads();
        lo = s_lohi[0]; hi = s_lohi[1];
        const bool staged = tv == nullptr && hi >= lo && hi - lo < (uint32_t)SETUP_WIN && hi < p.nv;
        if (staged) {
            const uint8_t* base = reinterpret_cast<const uint8_t*>(verts);
            const size_t vbytes = (size_t)p.vwords * 4;
            const size_t b0 = (size_t)lo * vbytes, a0 = b0 & ~(size_t)15, b1 = (size_t)(hi + 1) * vbytes;
            const uint32_t n16 = (uint32_t)((b1 - a0 + 15) >> 4);
            for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) cp_async16(s_stage + (size_t)i * 16, base + a0 + (size_t)i * 16);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            const float* s_vert = reinterpret_cast<const float*>(s_stage + (b0 - a0));
            setup_group<true, SPOT>(group, verts, fc, s_vert, lo, tv, tex, lights, recs, keys, heads, s_mask, wire, st, p, n_op, n_tr, s_ord);
        } else {
            setup_group<false, SPOT>(group, verts, fc, nullptr, 0, tv, tex, lights, recs, keys, heads, s_mask, wire, st, p, n_op, n_tr, s_ord);
        }
        __syncthreads();
        masks_flush(s_mask, masks, n_mtiles, group, p, make_uint4(s_ord[0], s_ord[1], s_ord[2], s_ord[3]), ocount, st);
        __syncthreads();                                   // the masks are reused by the next group of a persistent CTA
    }
    // one pair of global atomics per block
    n_op = __reduce_add_sync(0xFFFFFFFFu, n_op); n_tr = __reduce_add_sync(0xFFFFFFFFu, n_tr);      // REDUX: one instruction each
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { if (n_op) atomicAdd(&s_cnt[0], n_op); if (n_tr) atomicAdd(&s_cnt[1], n_tr); }
    __syncthreads();
    if (threadIdx.x == 0) { if (s_cnt[0]) atomicAdd(&st->n_opaque, s_cnt[0]); if (s_cnt[1]) atomicAdd(&st->n_transp, s_cnt[1]); }
    host_signal_done(p, st, HS_SETUP);
}

// =================================================================================================
// fragment evaluation shared by both fill kernels (render.rs:1534-1661)
// =================================================================================================
struct Pixel { uint32_t rgba; float z; };

// Edge functions + barycentrics at pixel (x,y): render.rs:1517-1542, 1706-1712.
template <typename Rec>
__device__ __forceinline__ bool inside_test(const Rec& r, uint32_t x, uint32_t y, float& bc_x, float& bc_y, float& bc_z) {
    uint32_t min_x = r.bbox_x & 0xFFFF, min_y = r.bbox_y & 0xFFFF;
    float w0, w1;
    if (r.flags & SF_FAST_EDGE) {
        // all terms are integers below 2^23: every rounded add of the reference is exact, so the
        // stepped value equals the closed form
        float dx = (float)(x - min_x), dy = (float)(y - min_y);
        w0 = r.w0s + dy * r.b0 + dx * r.a0;
        w1 = r.w1s + dy * r.b1 + dx * r.a1;
    } else {
        // replay the reference's rounded additions: (y-min_y) row steps, then (x-min_x) pixel steps
        // (two independent dependent-add chains: unrolled so that the adds, not the loop control, fill the issue slots)
        w0 = r.w0s; w1 = r.w1s;
        const float b0 = r.b0, b1 = r.b1, a0 = r.a0, a1 = r.a1;
        uint32_t n = y > min_y ? y - min_y : 0u;
        for (; n >= 8; n -= 8) {
            #pragma unroll
            for (int k = 0; k < 8; ++k) { w0 = __fadd_rn(w0, b0); w1 = __fadd_rn(w1, b1); }
        }
        for (; n; --n) { w0 = __fadd_rn(w0, b0); w1 = __fadd_rn(w1, b1); }
        n = x > min_x ? x - min_x : 0u;
        for (; n >= 8; n -= 8) {
            #pragma unroll
            for (int k = 0; k < 8; ++k) { w0 = __fadd_rn(w0, a0); w1 = __fadd_rn(w1, a1); }
        }
        for (; n; --n) { w0 = __fadd_rn(w0, a0); w1 = __fadd_rn(w1, a1); }
    }
    bc_x = w0 * r.inv_area;
    bc_y = w1 * r.inv_area;
    bc_z = 1.0f - bc_x - bc_y;
    const float ERR = -0.0001f;
    return bc_x >= ERR && bc_y >= ERR && bc_z >= ERR;                              // :1541-1542
}

// n rounded additions of (s0, s1) to (w0, w1): the reference's incremental edge stepping (render.rs:1706-1712), unrolled
__device__ __forceinline__ void edge_steps(float& w0, float& w1, float s0, float s1, uint32_t n) {
    for (; n >= 8; n -= 8) {
        #pragma unroll
        for (int k = 0; k < 8; ++k) { w0 = __fadd_rn(w0, s0); w1 = __fadd_rn(w1, s1); }
    }
    for (; n; --n) { w0 = __fadd_rn(w0, s0); w1 = __fadd_rn(w1, s1); }
}

// inside_test for a pixel of a warp's block when the edge values of the block's row start (column max(bx0, min_x), this
// pixel's row) are already known (pw0, pw1: see k_fill_opaque, "shared edge prefix"): the same rounded additions as
// inside_test replays, minus the prefix the pixels of a block row have in common.  At most BW - 1 steps are left.
template <int BW, typename Rec>
__device__ __forceinline__ bool inside_test_prefix(const Rec& r, uint32_t x, uint32_t y, float pw0, float pw1, uint32_t bx0,
                                                   float& bc_x, float& bc_y, float& bc_z) {
    if (r.flags & SF_FAST_EDGE) return inside_test(r, x, y, bc_x, bc_y, bc_z);
    const uint32_t min_x = r.bbox_x & 0xFFFF;
    const uint32_t n = x - max(bx0, min_x);                 // the caller's bounding-box test guarantees x >= min_x (and x >= bx0)
    float w0 = pw0, w1 = pw1;
    const float a0 = r.a0, a1 = r.a1;
    #pragma unroll
    for (int k = 0; k < BW - 1; ++k) if ((uint32_t)k < n) { w0 = __fadd_rn(w0, a0); w1 = __fadd_rn(w1, a1); }
    bc_x = w0 * r.inv_area;
    bc_y = w1 * r.inv_area;
    bc_z = 1.0f - bc_x - bc_y;
    const float ERR = -0.0001f;
    return bc_x >= ERR && bc_y >= ERR && bc_z >= ERR;
}

// texel of the surface at barycentric (bc): render.rs:1563-1586, types.rs:671-681.  Returns its index
// in the texel pool, or TEXEL_NONE for a zero-sized texture (sample() = TRANSPARENT).
constexpr uint32_t TEXEL_NONE = 0xFFFFFFFFu;
template <typename Rec>
__device__ __forceinline__ uint32_t texel_index(const Rec& r, float bc_x, float bc_y, float bc_z, float inv_z,
                                                const TexDev& t, const CallParams& p) {
    float u, v;
    if (p.affine_textures) {                                                   // :1563-1567
        u = bc_x * r.u1 + bc_y * r.u2 + bc_z * r.u3;
        v = bc_x * r.v1 + bc_y * r.v2 + bc_z * r.v3;
    } else {                                                                   // :1568-1578
        float uo = bc_x * r.u1 * r.iz1 + bc_y * r.u2 * r.iz2 + bc_z * r.u3 * r.iz3;
        float vo = bc_x * r.v1 * r.iz1 + bc_y * r.v2 * r.iz2 + bc_z * r.v3 * r.iz3;
        u = uo / inv_z;
        v = vo / inv_z;
    }
    if (t.w == 0 || t.h == 0) return TEXEL_NONE;                               // types.rs:673-675
    float uw = rem_euclid1(u), vw = rem_euclid1(1.0f - v);                     // :1583, types.rs:676-677
    uint32_t tx = min(f2u32sat(uw * (float)t.w), t.w - 1);
    uint32_t ty = min(f2u32sat(vw * (float)t.h), t.h - 1);
    return t.off + ty * t.w + tx;
}

// The sampled texel of a fragment as one word: RGB555 path = the Color15 (0x7FFF = WHITE for untextured surfaces, 0x0000 =
// TRANSPARENT for zero-sized textures); RGB888 path = r | g<<8 | b<<16 | blend<<24 (WHITE/Opaque, or Erase for zero-sized).
template <bool RGB888>
__device__ __forceinline__ uint32_t sample_texel(const SurfRec& r, float bc_x, float bc_y, float bc_z, float inv_z,
                                                 const TexDev* __restrict__ tex, const void* __restrict__ texels, const CallParams& p) {
    if (!(r.flags & SF_TEXTURED)) return RGB888 ? 0x00FFFFFFu : 0x7FFFu;
    uint32_t ti = texel_index(r, bc_x, bc_y, bc_z, inv_z, tex[r.flags >> 16], p);
    if (ti == TEXEL_NONE) return RGB888 ? ((uint32_t)B32_BLEND_ERASE << 24) : 0u;
    return RGB888 ? __ldg(reinterpret_cast<const uint32_t*>(texels) + ti) : (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(texels) + ti);
}

// Transparency rules + colour pipeline of rasterize_triangle_15 (render.rs:1591-1661) on a sampled Color15.  Returns
// false when the texel is skipped; otherwise rgb = Color15::r8/g8/b8 of the final colour, semi = its bit 15.
__device__ __forceinline__ bool shade_color(const SurfRec& r, uint32_t x, uint32_t y, float bc_x, float bc_y, float bc_z, uint32_t color,
                                            const CallParams& p, uint32_t& o_r, uint32_t& o_g, uint32_t& o_b, bool& semi) {
    bool is_black = (color & 0x7FFF) == 0;                                         // :1591-1607
    if (color == 0) {
        if (!(r.flags & SF_BLACK_TR)) color = 0x8000; else return false;
    } else if ((r.flags & SF_BLACK_TR) && is_black) {
        return false;
    }
    uint32_t tr8 = expand5((color >> 10) & 31), tg8 = expand5((color >> 5) & 31), tb8 = expand5(color & 31);
    uint32_t vr = f2u8(bc_x * (float)(r.vc1 & 0xFF) + bc_y * (float)(r.vc2 & 0xFF) + bc_z * (float)(r.vc3 & 0xFF));          // :1618-1620
    uint32_t vg = f2u8(bc_x * (float)((r.vc1 >> 8) & 0xFF) + bc_y * (float)((r.vc2 >> 8) & 0xFF) + bc_z * (float)((r.vc3 >> 8) & 0xFF));
    uint32_t vb = f2u8(bc_x * (float)(r.vc1 >> 16) + bc_y * (float)(r.vc2 >> 16) + bc_z * (float)(r.vc3 >> 16));
    uint32_t mr = min((tr8 * vr) >> 7, 255u), mg = min((tg8 * vg) >> 7, 255u), mb = min((tb8 * vb) >> 7, 255u);   // :1624-1626
    float sr, sg, sb;                                                              // :1629-1640
    if (p.shading == B32_SHADE_NONE) { sr = sg = sb = 1.0f; }
    else if (p.shading == B32_SHADE_FLAT) { sr = r.sh[0]; sg = r.sh[1]; sb = r.sh[2]; }
    else {
        sr = bc_x * r.sh[0] + bc_y * r.sh[3] + bc_z * r.sh[6];
        sg = bc_x * r.sh[1] + bc_y * r.sh[4] + bc_z * r.sh[7];
        sb = bc_x * r.sh[2] + bc_y * r.sh[5] + bc_z * r.sh[8];
    }
    uint32_t r8 = f2u8(fminf((float)mr * rclamp(sr, 0.0f, 2.0f), 255.0f));          // :1643-1645
    uint32_t g8 = f2u8(fminf((float)mg * rclamp(sg, 0.0f, 2.0f), 255.0f));
    uint32_t b8 = f2u8(fminf((float)mb * rclamp(sb, 0.0f, 2.0f), 255.0f));
    uint32_t r5, g5, b5;
    if (r.flags & SF_DITHER) {                                                     // :1173-1182
        // PS1_DITHER_MATRIX rows packed as signed nibbles, :1150-1155
        const uint32_t rows = (y & 2) ? ((y & 1) ? 0xE2F3u : 0x0C1Du) : ((y & 1) ? 0xF3E2u : 0x1D0Cu);
        int32_t off = (int32_t)((rows >> ((x & 3) * 4)) & 0xF);
        off = (off ^ 8) - 8;                                                       // sign-extend the nibble
        r5 = (uint32_t)min(max(((int32_t)r8 + off) >> 3, 0), 31);
        g5 = (uint32_t)min(max(((int32_t)g8 + off) >> 3, 0), 31);
        b5 = (uint32_t)min(max(((int32_t)b8 + off) >> 3, 0), 31);
    } else {
        r5 = r8 >> 3; g5 = g8 >> 3; b5 = b8 >> 3;
    }
    semi = (color & 0x8000) || (r5 == 0 && g5 == 0 && b5 == 0);                    // :1659-1661
    o_r = expand5(r5); o_g = expand5(g5); o_b = expand5(b5);                       // Color15::r8/g8/b8
    return true;
}

// Texture sample (render.rs:1563-1586) + shade_color.
__device__ __forceinline__ bool shade(const SurfRec& r, uint32_t x, uint32_t y, float bc_x, float bc_y, float bc_z, float inv_z,
                                      const TexDev* __restrict__ tex, const uint16_t* __restrict__ texels, const CallParams& p,
                                      uint32_t& o_r, uint32_t& o_g, uint32_t& o_b, bool& semi) {
    return shade_color(r, x, y, bc_x, bc_y, bc_z, sample_texel<false>(r, bc_x, bc_y, bc_z, inv_z, tex, texels, p), p, o_r, o_g, o_b, semi);
}

// RGB888 colour pipeline of rasterize_triangle (render.rs:1350-1389) on a sampled Color word: Erase texels skip,
// Color::modulate (types.rs:801-808), shade_color_rgb WITHOUT a clamp of the factor (render.rs:1074-1081),
// apply_dither re-expanded with `<< 3` (render.rs:1186-1197).  blend = the texel's tag.
__device__ __forceinline__ bool shade_color888(const SurfRec& r, uint32_t x, uint32_t y, float bc_x, float bc_y, float bc_z, uint32_t c,
                                               const CallParams& p, uint32_t& o_r, uint32_t& o_g, uint32_t& o_b, uint32_t& o_blend) {
    uint32_t cr = c & 0xFF, cg = (c >> 8) & 0xFF, cb = (c >> 16) & 0xFF, blend = c >> 24;
    if (blend == B32_BLEND_ERASE) return false;                                    // :1350-1354
    uint32_t vr = f2u8(bc_x * (float)(r.vc1 & 0xFF) + bc_y * (float)(r.vc2 & 0xFF) + bc_z * (float)(r.vc3 & 0xFF));          // :1357-1362
    uint32_t vg = f2u8(bc_x * (float)((r.vc1 >> 8) & 0xFF) + bc_y * (float)((r.vc2 >> 8) & 0xFF) + bc_z * (float)((r.vc3 >> 8) & 0xFF));
    uint32_t vb = f2u8(bc_x * (float)(r.vc1 >> 16) + bc_y * (float)(r.vc2 >> 16) + bc_z * (float)(r.vc3 >> 16));
    uint32_t mr = min((cr * vr) >> 7, 255u), mg = min((cg * vg) >> 7, 255u), mb = min((cb * vb) >> 7, 255u);   // :1365
    float sr, sg, sb;                                                              // :1368-1381
    if (p.shading == B32_SHADE_NONE) { sr = sg = sb = 1.0f; }
    else if (p.shading == B32_SHADE_FLAT) { sr = r.sh[0]; sg = r.sh[1]; sb = r.sh[2]; }
    else {
        sr = bc_x * r.sh[0] + bc_y * r.sh[3] + bc_z * r.sh[6];
        sg = bc_x * r.sh[1] + bc_y * r.sh[4] + bc_z * r.sh[7];
        sb = bc_x * r.sh[2] + bc_y * r.sh[5] + bc_z * r.sh[8];
    }
    uint32_t r8 = f2u8(fminf((float)mr * sr, 255.0f));                             // :1383
    uint32_t g8 = f2u8(fminf((float)mg * sg, 255.0f));
    uint32_t b8 = f2u8(fminf((float)mb * sb, 255.0f));
    if (r.flags & SF_DITHER) {                                                     // :1387-1389
        const uint32_t rows = (y & 2) ? ((y & 1) ? 0xE2F3u : 0x0C1Du) : ((y & 1) ? 0xF3E2u : 0x1D0Cu);
        int32_t off = (int32_t)((rows >> ((x & 3) * 4)) & 0xF);
        off = (off ^ 8) - 8;
        r8 = (uint32_t)min(max(((int32_t)r8 + off) >> 3, 0), 31) << 3;
        g8 = (uint32_t)min(max(((int32_t)g8 + off) >> 3, 0), 31) << 3;
        b8 = (uint32_t)min(max(((int32_t)b8 + off) >> 3, 0), 31) << 3;
    }
    o_r = r8; o_g = g8; o_b = b8; o_blend = blend;
    return true;
}

// Texture::sample (types.rs:1242-1253) + shade_color888.
__device__ __forceinline__ bool shade888(const SurfRec& r, uint32_t x, uint32_t y, float bc_x, float bc_y, float bc_z, float inv_z,
                                         const TexDev* __restrict__ tex, const uint32_t* __restrict__ texels, const CallParams& p,
                                         uint32_t& o_r, uint32_t& o_g, uint32_t& o_b, uint32_t& o_blend) {
    return shade_color888(r, x, y, bc_x, bc_y, bc_z, sample_texel<true>(r, bc_x, bc_y, bc_z, inv_z, tex, texels, p), p, o_r, o_g, o_b, o_blend);
}

// =================================================================================================
// k_fill_opaque — pass 1, order-free, visibility first
// =================================================================================================
// One CTA per 16x16 screen tile (all CTAs of a 320x240 frame are resident at once); one
// warp per 4x4 pixel block; a pixel is owned by TWO lanes (lane and lane+16) that evaluate different
// surfaces at the same time and merge their winners — the winner rule is associative, so the merge is
// exact.
//   1. the tile's bin is copied in walk-key order with one counting-sort pass — into shared memory up to
//      OP_SORT_MAX entries, into a global scratch beyond: painter's mode = nearest (last drawn) first, z-buffer
//      mode = smallest depth lower bound first.  The order is an efficiency device only: the per-pixel winner
//      rule is exact for ANY order, ties and all.
//   2. the CTA streams the visibility part (80 of 128 bytes) of the surface records of the walk order through a
//      3-deep shared-memory ring, OP_CHUNK records per step, in 16-byte cp.async pieces: the loads of steps c+1 and
//      c+2 are in flight while the warps work on step c, so no warp ever waits for an L2 round trip of its own.
//   3. every warp filters the step's entries one per lane (bbox vs the block's still-open pixels,
//      priority / depth bound vs the block's weakest pixel); survivors are evaluated two per half-warp at
//      a time.  The walk only decides WHO wins each pixel: inside test, depth, and — for black-keyed
//      textured surfaces — whether the texel writes at all, answered by a 1-bit-per-texel mask of the
//      texel pool that a TMA bulk copy (cp.async.bulk + mbarrier) put into shared memory at CTA start
//      (8 KB for a 256x256 atlas; larger pools read the mask through L1).
//   4. a warp stops as soon as no later entry can change any pixel of its block; the CTA stops when all
//      its warps have.
//   5. each pixel shades its winner once (texture, modulate, lighting, dither), at the end.
#ifndef B32_OP_THREADS
#define B32_OP_THREADS 512        // one CTA per 16x16 tile: the tile's sort and record staging are done once, not per half
#endif
#ifndef B32_OP_DUAL
#define B32_OP_DUAL 1
#endif
#ifndef B32_OP_CHUNK
#define B32_OP_CHUNK 128
#endif
// The kernel exists in two shapes, picked per call by the number of tiles (launch_fill_opaque):
//   OpDense  512 threads, a warp owns 4x4 pixels with TWO lanes per pixel (two surfaces in flight per pixel): more
//            instruction-level parallelism per tile; 3 CTAs per SM = 444 tiles in one wave.  Best for a blocking call
//            whose tiles fit that wave (320x240 = 300 tiles): C4 24.5 vs 26.6 us.
//   OpSparse 256 threads, a warp owns 8x4 pixels, one lane per pixel; ring steps of 96 records keep a CTA at 53 KB of shared
//            memory: 4 CTAs per SM = 592 tiles per wave.  Best for larger framebuffers, where the tiles come in several
//            waves of latency-bound CTAs (sample levels at 640x480 20.6 vs 28.7 us, C4 at 640x480 39.9 vs 49.2 us, at
//            1920x1080 144 vs 213 us), and for enqueued frames that overlap with their neighbours on the GPU.
template <int THREADS_, bool DUAL_, int CHUNK_, int MINB_>
struct OpCfg {
    static constexpr int THREADS = THREADS_;
    static constexpr bool DUAL = DUAL_;          // true: a warp owns 4x4 pixels, two lanes per pixel; false: 8x4 pixels, one lane per pixel
    static constexpr int WARPS = THREADS / 32;
    static constexpr int BW = DUAL ? 4 : 8, BH = 4;                      // pixel block of one warp
    static constexpr int WPT = (TILE_W / BW) * (TILE_H / BH);            // warps per 16x16 tile
    static constexpr int SPLIT = WPT / WARPS;    // CTAs per tile (256 threads, dual: 2 = half tiles of 16x8 px)
    static_assert(SPLIT >= 1 && SPLIT * WARPS == WPT, "a CTA covers a whole number of block rows of one tile");
    static constexpr int CHUNK = CHUNK_;         // surface records (their 80-byte visibility part) staged per step: most warps are done within
                                                 // the first step, so the CTA-wide barrier between steps rarely holds anybody up
    static constexpr int REC_PIECES = sizeof(SurfHot) / 16;
    static_assert(CHUNK % 32 == 0 && CHUNK <= 256, "whole 32-entry batches; slots are stored in a byte");
    static_assert((size_t)3 * CHUNK * sizeof(SurfHot) >= (size_t)OP_SORT_MAX_ENTRIES * (sizeof(BinHead) + 4), "the ring area doubles as the candidate list + the heads of a depth window");
    static constexpr int BUCKETS = THREADS < 256 ? THREADS : 256;        // key buckets of the counting sort (one scan thread each)
    static constexpr int BUCKET_BITS = BUCKETS == 256 ? 8 : (BUCKETS == 128 ? 7 : 6);
    static constexpr int RING = 3;               // ring depth: steps c, c+1, c+2
    static constexpr int SORT_MAX = OP_SORT_MAX_ENTRIES;   // bin entries orderable in shared memory (16 KB of heads)
    static constexpr int TEX_SMEM = 256;         // texture descriptors cached in shared memory
    static constexpr size_t SMEM = (size_t)SORT_MAX * sizeof(BinHead) + (size_t)RING * CHUNK * sizeof(SurfHot) +
                                   (size_t)TEX_SMEM * sizeof(TexDev) + (size_t)OP_MASK_SMEM_WORDS * 4 +
                                   (size_t)WARPS * 32 * sizeof(uint2) + (size_t)WARPS * 32;
    static constexpr int MINB = MINB_;           // CTAs per SM the register allocation must allow (shared memory permitting)
};
#ifndef B32_OP_MINB
#define B32_OP_MINB (B32_OP_THREADS == 512 ? 3 : 5)   // 512 threads: 3 CTAs per SM (444 slots >= the 300 tiles of a 320x240 frame: one wave)
#endif
#ifndef B32_OPS_CHUNK
#define B32_OPS_CHUNK 96          // 53 KB of shared memory per CTA: 4 CTAs per SM
#endif
#ifndef B32_OPS_MINB
#define B32_OPS_MINB 4
#endif
using OpDense = OpCfg<B32_OP_THREADS, B32_OP_DUAL != 0, B32_OP_CHUNK, B32_OP_MINB>;
using OpSparse = OpCfg<256, false, B32_OPS_CHUNK, B32_OPS_MINB>;
#ifdef B32_FILL_STATS
__device__ uint32_t g_fill_stats[4096 * 16 * 8];     // [tile][warp][8]: t_start, t_sorted, t_end, batches, t_first_data, t_batch0_end, t_batch1_end, t_loop_end
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t gtime() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (uint32_t)t; }
#endif

// ---- crowded tiles: depth-ordered windows out of a scratch slice ----------------------------------------------------
// A tile with more candidates than one window holds (a million-triangle mesh puts thousands of surfaces behind every
// tile) reserves a slice of a global scratch with ONE atomic, writes the heads of all its pass-1 candidates there (one
// pass over its mask row), histograms their walk keys into `buckets` key buckets, and then takes its windows in DEPTH
// order — the nearest few buckets first — so that the early-out of one window holds for every later one and the tile
// stops as soon as every pixel is settled, usually after its first window.  A single bucket that holds more than a
// window (many equal keys) is taken in slice order, a window at a time, without that carry-over.  If the scratch is
// exhausted the tile falls back to windows in list order.  All state lives in shared memory; these functions are
// deliberately not inlined (they run for a handful of tiles of unusual frames and must not cost the usual tile anything).
struct CrowdShared {
    uint32_t n_valid, kmin, shift, b_next, over_b, over_i, item[4], ncol, lo, hi;
};

// slot for every lane with `take` set: one shared-memory atomic per warp (all 32 lanes must call)
__device__ __forceinline__ uint32_t warp_append(bool take, uint32_t* counter) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, take), lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0 && m) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    return base + __popc(m & ((1u << lane) - 1));
}

__device__ __forceinline__ uint32_t crowd_bucket(uint32_t k, uint32_t kmin, uint32_t shift, uint32_t buckets) {
    return k == 0xFFFFFFFFu ? 0u : (buckets - 1) - ((k - kmin) >> shift);
}

// pass 1: every candidate's head -> slice[] (those that are pass-1 surfaces touching this tile), key range, histogram.
// The face indices of all candidates are written once to ids[] (= the tail of the tile's scratch slice),
// with the mask loads batched four at a time; the heads are then gathered with coalesced index reads.
__device__ __noinline__ void crowd_prepare(const uint4* __restrict__ mrow, const CandScan cs, uint32_t n_cand, const BinHead* __restrict__ heads,
                                           uint32_t tpx0, uint32_t tpy0, BinHead* __restrict__ slice, uint32_t* __restrict__ ids,
                                           uint32_t* s_ghist, uint32_t buckets, uint32_t bucket_bits, CrowdShared* cr) {
    if (threadIdx.x == 0) { cr->ncol = 0; cr->lo = 0xFFFFFFFFu; cr->hi = 0; cr->b_next = 0; cr->over_b = 0xFFFFFFFFu; cr->over_i = 0; }
    for (uint32_t i = threadIdx.x; i < buckets; i += blockDim.x) s_ghist[i] = 0;
    uint32_t idx = cs.base;
    for (uint32_t g = threadIdx.x; g < cs.n_groups; g += 4 * cs.stride) {
        uint4 m[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = g + j * cs.stride < cs.n_groups ? mrow[g + j * cs.stride] : make_uint4(0, 0, 0, 0);
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            if ((m[j].x | m[j].y | m[j].z | m[j].w) == 0) continue;
            const uint32_t words[4] = {m[j].x, m[j].y, m[j].z, m[j].w};
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t bits = words[q];
                while (bits) { const uint32_t bb = __ffs(bits) - 1; bits &= bits - 1; ids[idx++] = (g + j * cs.stride) * SETUP_GROUP + q * 32 + bb; }
            }
        }
    }
    __syncthreads();                                       // ids[] written by this block: visible to it after the barrier
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (uint32_t w0 = threadIdx.x & ~31u; w0 < n_cand; w0 += 4 * blockDim.x) {       // warp-uniform trip count
        BinHead h[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) { const uint32_t i = w0 + (threadIdx.x & 31) + j * blockDim.x; if (i < n_cand) h[j] = heads[ids[i]]; else h[j].bbox_x = 0; }
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t min_x = h[j].bbox_x & 0xFFFF, max_x = h[j].bbox_x >> 16, min_y = h[j].bbox_y & 0xFFFF, max_y = h[j].bbox_y >> 16;
            const bool ok = !(h[j].bbox_x == 0 || max_x <= tpx0 || min_x >= tpx0 + TILE_W || max_y <= tpy0 || min_y >= tpy0 + TILE_H);
            const uint32_t pos = warp_append(ok, &cr->ncol);
            if (ok) {
                slice[pos] = h[j];
                if (h[j].key != 0xFFFFFFFFu) { lo = min(lo, h[j].key); hi = max(hi, h[j].key); }
            }
        }
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo); hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(&cr->lo, lo); atomicMax(&cr->hi, hi); }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t kmin = cr->lo, kmax = cr->hi;
        if (kmin > kmax) { kmin = 0; kmax = 0; }
        const uint32_t range = kmax - kmin;
        cr->kmin = kmin;
        cr->shift = range >= buckets ? (32 - __clz(range)) - bucket_bits : 0;
        cr->n_valid = cr->ncol;
    }
    __syncthreads();
    const uint32_t nv = cr->n_valid, kmin = cr->kmin, shift = cr->shift;
    for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) atomicAdd(&s_ghist[crowd_bucket(slice[i].key, kmin, shift, buckets)], 1u);   // (this thread block wrote slice[]: visible after the barrier)
    __syncthreads();
}

// the next depth window: its heads -> s_tmp[0..n).  Returns n; 0xFFFFFFFF = no window left.  *carry = the window is one
// of the depth sequence (its early-outs hold for every later window).
__device__ __noinline__ uint32_t crowd_next_window(const BinHead* __restrict__ slice, uint32_t tile_weak, uint32_t* s_ghist, uint32_t buckets,
                                                   uint32_t win, BinHead* s_tmp, CrowdShared* cr, bool* carry) {
    __syncthreads();                                       // the previous window's readers of cr->item are through
    if (threadIdx.x == 0) {
        cr->ncol = 0;
        if (cr->over_b != 0xFFFFFFFFu) {                   // an over-full bucket, the next `win` slice entries
            cr->item[0] = cr->over_b; cr->item[1] = cr->over_b + 1; cr->item[2] = cr->over_i; cr->item[3] = min(cr->over_i + win, cr->n_valid);
            cr->over_i += win;
            if (cr->over_i >= cr->n_valid) cr->over_b = 0xFFFFFFFFu;
        } else {
            uint32_t b = cr->b_next;
            while (b < buckets && s_ghist[b] == 0) ++b;
            if (b >= buckets) { cr->item[0] = 0xFFFFFFFFu; }
            else if (s_ghist[b] > win) {                   // over-full: this call takes its first `win` slice entries
                cr->item[0] = b; cr->item[1] = b + 1; cr->item[2] = 0; cr->item[3] = min(win, cr->n_valid);
                cr->b_next = b + 1;
                if (win < cr->n_valid) { cr->over_b = b; cr->over_i = win; }
            } else {
                uint32_t sum = 0, first = b;
                while (b < buckets && sum + s_ghist[b] <= win) sum += s_ghist[b++];
                cr->item[0] = first; cr->item[1] = b; cr->item[2] = 0; cr->item[3] = cr->n_valid;
                cr->b_next = b;
            }
        }
    }
    __syncthreads();
    const uint32_t b0 = cr->item[0], b1 = cr->item[1], i0 = cr->item[2], i1 = cr->item[3];
    if (b0 == 0xFFFFFFFFu) return 0xFFFFFFFFu;
    *carry = (i0 == 0 && i1 == cr->n_valid && cr->over_b == 0xFFFFFFFFu);
    const uint32_t kmin = cr->kmin, shift = cr->shift;
    for (uint32_t w = i0 + (threadIdx.x & ~31u); w < i1; w += blockDim.x) {             // warp-uniform trip count
        const uint32_t i = w + (threadIdx.x & 31);
        BinHead h{0, 0, 0, 0};
        bool take = false;
        if (i < i1) {
            h = slice[i];
            const uint32_t b = crowd_bucket(h.key, kmin, shift, buckets);
            // (z-buffer: key 0xFFFFFFFF = "no bound claimed" always passes; equal bounds pass: ties are resolved per pixel)
            take = b >= b0 && b < b1 && h.key >= tile_weak;
        }
        const uint32_t pos = warp_append(take, &cr->ncol);
        if (take && pos < win) s_tmp[pos] = h;
    }
    __syncthreads();
    return min(cr->ncol, win);
}

// RGB888 = the render_mesh instantiation (8-bit colour pipeline in step 5; everything else is shared); C = OpDense / OpSparse.
// PRE = shared edge prefix: calls whose surfaces replay the reference's rounded edge additions (float / ortho projection:
// no SF_FAST_EDGE) compute, per 8 survivors, the edge values at the start of each of the block's 4 rows ONCE — one
// (survivor, row) pair per lane, 32 chains side by side — and every pixel only adds its last < BW steps, instead of every
// lane replaying the whole O(bbox width + height) chain of every survivor.  Same additions in the same order: bit-exact.
template <bool RGB888, class C, bool PRE>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
k_fill_opaque(const SurfRec* __restrict__ recs, const uint4* __restrict__ masks, const BinHead* __restrict__ heads,
              const TexDev* __restrict__ tex, const uint16_t* __restrict__ texels, const uint32_t* __restrict__ texmask,
              uint32_t* __restrict__ fb_rgba, float* __restrict__ fb_z, CallState* __restrict__ st,
              uint32_t* __restrict__ sticky, BinHead* __restrict__ crowd, uint32_t crowd_cap, CallParams p) {
    constexpr int OP_THREADS = C::THREADS, OP_WARPS = C::WARPS, OP_BW = C::BW, OP_BH = C::BH, OP_SPLIT = C::SPLIT;
    constexpr bool OP_DUAL = C::DUAL;
    constexpr int OP_CHUNK = C::CHUNK, OP_REC_PIECES = C::REC_PIECES, OP_BUCKETS = C::BUCKETS, OP_BUCKET_BITS = C::BUCKET_BITS;
    constexpr int OP_RING = C::RING, OP_SORT_MAX = C::SORT_MAX, OP_TEX_SMEM = C::TEX_SMEM;
    extern __shared__ __align__(128) uint8_t op_smem[];
    SurfHot* s_rec = reinterpret_cast<SurfHot*>(op_smem);                               // [OP_RING][OP_CHUNK] record ring (visibility part)
    BinHead* s_sh = reinterpret_cast<BinHead*>(s_rec + OP_RING * OP_CHUNK);             // [OP_SORT_MAX] the window's heads in walk order
    TexDev* s_tex = reinterpret_cast<TexDev*>(s_sh + OP_SORT_MAX);                      // [OP_TEX_SMEM]
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_tex + OP_TEX_SMEM);                // [OP_MASK_SMEM_WORDS] "texel writes" bits
    uint2* s_surv = reinterpret_cast<uint2*>(s_mask + OP_MASK_SMEM_WORDS);              // [OP_WARPS][32] survivors: (key, face)
    uint8_t* s_sidx = reinterpret_cast<uint8_t*>(s_surv + OP_WARPS * 32);               // [OP_WARPS][32] ... and their slot in the ring step
    uint32_t* s_cand = reinterpret_cast<uint32_t*>(s_rec);                              // [OP_SORT_MAX] face indices of the window (the ring is idle then)
    __shared__ uint32_t s_hist[OP_BUCKETS];
    __shared__ uint32_t s_wsum[OP_THREADS / 32];
    __shared__ uint32_t s_minmax[2];
    __shared__ uint32_t s_nsmall, s_nraw;
    __shared__ uint32_t s_tile_weak;       // what the tile's weakest pixel still accepts after the windows done so far (see the window loop)
    __shared__ uint32_t s_ghist[OP_BUCKETS];   // crowded tiles: histogram of all candidates' walk keys
    __shared__ CrowdShared s_crowd;
    __shared__ uint32_t s_slice;
    __shared__ __align__(8) uint64_t s_mbar;
    if (threadIdx.x == 0) s_nraw = 0;
    // ---- prologue: nothing here reads what k_setup writes (the framebuffer included: the frame's clear may ride in
    //      k_setup), so it runs while k_setup finishes ----------------------------------------------------------------
    const uint32_t tile = blockIdx.x / OP_SPLIT, half = blockIdx.x % OP_SPLIT;
    // the "texel writes" mask of the whole texel pool travels by TMA while the candidates are collected
    const bool mask_staged = p.mask_smem_words != 0;
    if (mask_staged && threadIdx.x == 0) { mbar_init(&s_mbar, 1); bulk_g2s(s_mask, texmask, p.mask_smem_words * 4, &s_mbar); }
    const uint32_t* maskw = mask_staged ? s_mask : texmask;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t pix = OP_DUAL ? (lane & 15) : lane, sub = OP_DUAL ? (lane >> 4) : 0;
    const uint32_t tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    // thread -> pixel: each warp owns a block of its (half) tile; dual: lanes l and l+16 share a pixel
    constexpr uint32_t BPR = TILE_W / OP_BW;               // warp blocks per tile row
    const uint32_t wt = half * OP_WARPS + warp;            // this warp's block within the tile
    const uint32_t bx0 = tx * TILE_W + (wt % BPR) * OP_BW, by0 = ty * TILE_H + (wt / BPR) * OP_BH;
    const uint32_t x = bx0 + (pix % OP_BW), y = by0 + (pix / OP_BW);
    const bool valid = x < p.width && y < p.height;
    const bool tex_cached = p.ntex <= OP_TEX_SMEM;
    if (tex_cached) for (uint32_t i = threadIdx.x; i < p.ntex; i += OP_THREADS) s_tex[i] = tex[i];
    const TexDev* texd = tex_cached ? s_tex : tex;
    pdl_wait();                                            // k_setup has completed: records, heads, masks, counters, cleared framebuffer
    pdl_launch_dependents();                               // an enqueued ordered pass may be scheduled behind this grid (it waits for its completion)
    host_stamp_start(p, HS_FILL);
    // the pixel's framebuffer content
    Pixel px{0, 0.0f};
    if (valid) { px.rgba = fb_rgba[y * p.width + x]; px.z = fb_z[y * p.width + x]; }
    bool skip;
    {
        CallState s = *st;
        bool aborts = call_aborts(s, p.use_zbuffer, RGB888);
        if (p.async_call && blockIdx.x == 0 && threadIdx.x == 0) {                                // enqueue-only callers
            if (aborts) atomicOr(sticky, s.oob == 2 ? 32u : s.oob ? 1u : 2u);
            else if (s.n_transp && !p.enq_ordered) atomicOr(sticky, 8u);                         // pass 2 exists but was not enqueued
        }
        // large stepped surfaces in a fixed-point call (k_setup counted them): one 4-byte store into host-mapped memory, from
        // which the host picks the fill instantiation of this context's next calls (see launch_fill_opaque)
        if (blockIdx.x == 0 && threadIdx.x == 0 && s.n_big_stepped) *reinterpret_cast<volatile uint32_t*>(p.stepped_seq_host) = p.call_seq;
        // x-ray (render_mesh_15) and any framebuffer-reading surface (render_mesh) go through the ordered replay instead
        skip = aborts || (p.xray_mode && !RGB888) || (RGB888 && s.n_transp);
    }
    // ---- 0. this tile's candidates: the set bits of its mask row (see b32_device.cuh) ------------------------------
    // Usual case (the row fits OP_KREG masks per thread and the tile has at most OP_SORT_MAX candidates): every thread takes
    // the groups tid, tid + THREADS, ..., reserves room for its set bits with one shared-memory atomic and writes the face
    // indices — any order will do, the list is ordered by walk key next.  Otherwise: cand_scan / cand_expand, a window at a time.
    const uint32_t mtile = (ty >> p.mshift) * p.mtiles_x + (tx >> p.mshift);
    const uint4* mrow = masks + (size_t)mtile * p.n_groups;
    constexpr int OP_KREG = 4;
    const bool fast = !skip && p.n_groups <= (uint32_t)(OP_KREG * OP_THREADS);
    uint32_t n_cand = 0;
    bool fast_ok = false;
    __syncthreads();                                       // s_nraw = 0 is visible
    if (fast) {
        uint4 m[OP_KREG];
        uint32_t cnt = 0;
        #pragma unroll
        for (int j = 0; j < OP_KREG; ++j) {
            const uint32_t g = j * OP_THREADS + threadIdx.x;
            m[j] = g < p.n_groups ? mrow[g] : make_uint4(0, 0, 0, 0);
            cnt += popc128(m[j]);
        }
        uint32_t pos = cnt ? atomicAdd(&s_nraw, cnt) : 0u;
        if (pos + cnt <= (uint32_t)OP_SORT_MAX) {
            #pragma unroll
            for (int j = 0; j < OP_KREG; ++j) {
                if ((m[j].x | m[j].y | m[j].z | m[j].w) == 0) continue;       // most (tile, group) pairs are empty
                const uint32_t f0 = (j * OP_THREADS + threadIdx.x) * SETUP_GROUP;
                const uint32_t words[4] = {m[j].x, m[j].y, m[j].z, m[j].w};
                #pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t bits = words[q];
                    while (bits) { const uint32_t b = __ffs(bits) - 1; bits &= bits - 1; s_cand[pos++] = f0 + q * 32 + b; }
                }
            }
        }
        __syncthreads();
        n_cand = s_nraw;
        fast_ok = n_cand <= (uint32_t)OP_SORT_MAX;
    }
    CandScan cs{0, 0, 0, 0};
    if (!fast_ok) {
        cs = cand_scan<OP_THREADS>(mrow, skip ? 0u : p.n_groups, s_wsum);
        n_cand = cs.total;
    }
    if (n_cand == 0) {                                     // nothing to draw here; the mask copy must land before the CTA exits
        if (mask_staged && threadIdx.x == 0) while (!mbar_try_wait(&s_mbar, 0)) {}
        host_signal_done(p, st, HS_FILL);
        return;
    }
#ifdef B32_FILL_STATS
    uint32_t st_t0 = gtime(), st_batches = 0, st_surv = 0, st_inside = 0, st_shaded = 0;
    uint32_t st_t1 = 0, st_tl = 0, st_tfirst = 0, st_tb0 = 0, st_tb1 = 0;
#endif
    const Pixel px0 = px;
    // painter's: best = (key << 32 | face) + 1 of the winner so far (0 = framebuffer content)
    // z-buffer : best_face = face + 1 of the winner so far (0 = framebuffer content), its depth in px.z
    uint64_t best = valid ? 0ull : ~0ull;
    uint32_t best_face = 0;
    float win_bx = 0.0f, win_by = 0.0f;                    // PRE: the winner's barycentrics as the walk computed them (step 5 then needs no second replay of its chain)
    uint2* my_surv = s_surv + warp * 32;
    uint8_t* my_sidx = s_sidx + warp * 32;
    const bool offscreen = bx0 >= p.width || by0 >= p.height;      // whole warp off-screen
    const uint32_t tpx0 = tx * TILE_W, tpy0 = ty * TILE_H;
    bool mask_waited = !mask_staged;

    // The candidates are taken OP_SORT_MAX at a time (a "window"; almost always there is one).  The winner rule is
    // order-free, so windows need no order among themselves: each is ordered by walk key on its own (an efficiency
    // device for the early-out), and the winners carry over.
    // From the second window on, a candidate that cannot beat the tile's weakest pixel is dropped before it is ordered:
    // painter's: tile_weak = smallest winner key in the tile (0 while a pixel has no winner): keys below it lose everywhere;
    // z-buffer : tile_weak = ~bits(largest depth in the tile): a surface whose depth lower bound is behind it loses everywhere.
    uint32_t tile_weak = 0;
    // crowded tile: a slice of the global scratch for its heads (see crowd_prepare); none left = windows in list order
    BinHead* slice = nullptr;
    const uint32_t slice_len = n_cand + (n_cand + 3) / 4;          // n_cand heads, then n_cand face indices (4 per head-sized slot)
#ifndef B32_NO_CROWD
    if (n_cand > (uint32_t)OP_SORT_MAX) {
        if (threadIdx.x == 0) {
            uint32_t base = crowd_cap >= slice_len ? atomicAdd(&st->crowd_used, slice_len) : 0xFFFFFFFFu;
            s_slice = (base != 0xFFFFFFFFu && (uint64_t)base + slice_len <= crowd_cap) ? base : 0xFFFFFFFFu;
        }
        __syncthreads();
        if (s_slice != 0xFFFFFFFFu) slice = crowd + s_slice;
    }
#endif
    BinHead* s_tmp = reinterpret_cast<BinHead*>(s_cand + OP_SORT_MAX);         // [OP_SORT_MAX] a depth window's heads (the ring is idle then)
    if (slice) crowd_prepare(mrow, cs, n_cand, heads, tpx0, tpy0, slice, reinterpret_cast<uint32_t*>(slice + n_cand), s_ghist, OP_BUCKETS, OP_BUCKET_BITS, &s_crowd);
#ifdef B32_FILL_STATS
    uint32_t st_tprep = gtime(), st_twin = 0;
#endif
    bool gdone = offscreen;                                // depth-ordered windows: this warp's early-out, carried from window to window
    // One window.  MULTI = false_type is the usual tile (one window, no scratch slice, nothing carried over): the same body
    // with everything that only several windows need compiled out.  Returns false when no further window is needed.
    auto run_window = [&](auto multi_tag, const uint32_t win) -> bool {
        constexpr bool MULTI = decltype(multi_tag)::value;
        const uint32_t w0 = win * OP_SORT_MAX;
        if (!(MULTI && slice) && w0 >= n_cand) return false;
        const uint32_t wn = (MULTI && slice) ? 0u : min((uint32_t)OP_SORT_MAX, n_cand - w0);
        if (MULTI && win) {
            cp_async_wait<0>();
            if (threadIdx.x == 0) s_tile_weak = 0xFFFFFFFFu;
            __syncthreads();                               // the previous window's ring traffic is over: its area is reused
            uint32_t wk;
            if (!p.use_zbuffer) {
                bool allw = __all_sync(0xFFFFFFFFu, best != 0);                              // invalid lanes carry ~0
                wk = allw ? __reduce_min_sync(0xFFFFFFFFu, (uint32_t)((best - 1) >> 32)) : 0u;
                if (offscreen) wk = 0xFFFFFFFFu;
            } else {
                // ~bits of the largest depth (depths >= 0 here or the bound is not used): smaller = farther, as the walk keys
                float zmax = valid ? px.z : 0.0f;
                uint32_t zb = (zmax >= 0.0f) ? ~__float_as_uint(zmax) : 0u;                  // negative / NaN depths: no bound
                if (!valid) zb = 0xFFFFFFFFu;
                wk = __reduce_min_sync(0xFFFFFFFFu, zb);
            }
            if (lane == 0) atomicMin(&s_tile_weak, wk);
            __syncthreads();
            tile_weak = s_tile_weak;
            if (tile_weak == 0xFFFFFFFFu) tile_weak = 0;   // no on-screen pixel at all
        }
        // ---- 1. the window's heads (one 16-byte record per face, L2-resident) in registers; dropped: candidates that are not
        //         pass-1 surfaces (head bbox 0: pass 2) or, with coarse mask tiles, miss this 16x16 tile
        constexpr int KPT = OP_SORT_MAX / OP_THREADS;
        BinHead hh[KPT];
        uint32_t n = 0;
        bool carry = false;
        if (MULTI && slice) {
            n = crowd_next_window(slice, tile_weak, s_ghist, OP_BUCKETS, OP_SORT_MAX, s_tmp, &s_crowd, &carry);
#ifdef B32_FILL_STATS
            if (!win) st_twin = gtime();
#endif
            if (n == 0xFFFFFFFFu) return false;
            #pragma unroll
            for (int q = 0; q < KPT; ++q) {
                const uint32_t i = q * OP_THREADS + threadIdx.x;
                if (i < n) hh[q] = s_tmp[i]; else hh[q].bbox_x = 0;
            }
            __syncthreads();                               // s_tmp has been read: the ring may be written
        } else {
            if (!fast_ok) cand_expand(mrow, cs, w0, wn, s_cand);
            #pragma unroll
            for (int q = 0; q < KPT; ++q) {
                const uint32_t i = q * OP_THREADS + threadIdx.x;
                bool ok = false;
                if (i < wn) {
                    hh[q] = heads[s_cand[i]];
                    const uint32_t min_x = hh[q].bbox_x & 0xFFFF, max_x = hh[q].bbox_x >> 16, min_y = hh[q].bbox_y & 0xFFFF, max_y = hh[q].bbox_y >> 16;
                    ok = hh[q].bbox_x != 0 && !(max_x <= tpx0 || min_x >= tpx0 + TILE_W || max_y <= tpy0 || min_y >= tpy0 + TILE_H);
                    // (z-buffer: key 0xFFFFFFFF = "no bound claimed" always passes; equal bounds pass: ties are resolved per pixel)
                    if (MULTI) ok = ok && hh[q].key >= tile_weak;
                }
                if (!ok) hh[q].bbox_x = 0;
                n += __syncthreads_count(ok);              // (also orders the s_cand reads before the ring's writes)
            }
        }
        if (n == 0) return true;
        auto for_each_entry = [&](auto f) {
            #pragma unroll
            for (int q = 0; q < KPT; ++q) { if (hh[q].bbox_x) f(hh[q]); }
        };
        // ---- 1b. walk-key order, descending: one counting-sort pass into OP_BUCKETS key buckets (within a bucket the
        //          order is arbitrary; the early-out uses bucket bounds)
        uint32_t kmin = 0, shift = 0;
        const bool single_batch = n <= 32;                // one batch: nothing comes "later", so no order (and no early-out) is needed
        if (single_batch) {
            if (threadIdx.x == 0) s_nsmall = 0;
            __syncthreads();
            for_each_entry([&](const BinHead& h) { s_sh[atomicAdd(&s_nsmall, 1u)] = h; });
        } else {
            uint32_t lo = 0xFFFFFFFFu, hi = 0;
            for_each_entry([&](const BinHead& h) { if (h.key != 0xFFFFFFFFu) { lo = min(lo, h.key); hi = max(hi, h.key); } });   // 0xFFFFFFFF = "never cull": bucket 0
            if (threadIdx.x < OP_BUCKETS) s_hist[threadIdx.x] = 0;
            if (threadIdx.x == 0) { s_minmax[0] = 0xFFFFFFFFu; s_minmax[1] = 0; }
            lo = __reduce_min_sync(0xFFFFFFFFu, lo); hi = __reduce_max_sync(0xFFFFFFFFu, hi);
            __syncthreads();
            if (lane == 0) { atomicMin(&s_minmax[0], lo); atomicMax(&s_minmax[1], hi); }
            __syncthreads();
            kmin = s_minmax[0];
            uint32_t kmax = s_minmax[1];
            if (kmin > kmax) { kmin = 0; kmax = 0; }
            uint32_t range = kmax - kmin;
            shift = range >= OP_BUCKETS ? (32 - __clz(range)) - OP_BUCKET_BITS : 0;         // (range >> shift) <= OP_BUCKETS - 1
            // Bucket 0 belongs to the "never cull" keys ALONE: within a bucket the order is arbitrary, and the early-out
            // below takes the top of the current entry's bucket as the bound of everything still to come — a never-cull
            // entry behind a bounded one of the same bucket would be skipped by a bound that does not hold for it.  (Found
            // by the extended fuzz: ortho + z-buffer scenes, where both kinds of key occur, differed from run to run.)
            // The two highest key quanta therefore share bucket 1.
            auto bucket = [&](uint32_t k) { return k == 0xFFFFFFFFu ? 0u : (uint32_t)(OP_BUCKETS - 1) - min((k - kmin) >> shift, (uint32_t)(OP_BUCKETS - 2)); };
            for_each_entry([&](const BinHead& h) { atomicAdd(&s_hist[bucket(h.key)], 1u); });
            __syncthreads();
            uint32_t v = 0, xs = 0;                              // exclusive scan of the bucket counts (threads 0..OP_BUCKETS-1)
            if (threadIdx.x < OP_BUCKETS) {
                v = s_hist[threadIdx.x]; xs = v;
                for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, xs, o); if (lane >= (uint32_t)o) xs += t; }
                if (lane == 31) s_wsum[warp] = xs;
            }
            __syncthreads();
            if (threadIdx.x < OP_BUCKETS) {
                uint32_t pre = 0;
                for (uint32_t w = 0; w < warp; ++w) pre += s_wsum[w];
                s_hist[threadIdx.x] = pre + xs - v;
            }
            __syncthreads();
            for_each_entry([&](const BinHead& h) { s_sh[atomicAdd(&s_hist[bucket(h.key)], 1u)] = h; });
        }
        __syncthreads();                                      // the walk order, s_tex and the mbarrier are ready

        // ---- 2. the record ring: step c -> slot c % OP_RING; the 16-byte pieces of the step's records are dealt round robin
        auto stage = [&](uint32_t c) {
            for (uint32_t piece = threadIdx.x; piece < (uint32_t)(OP_CHUNK * OP_REC_PIECES); piece += OP_THREADS) {
                uint32_t slot = piece / OP_REC_PIECES, part = piece % OP_REC_PIECES;
                uint32_t e = c * OP_CHUNK + slot;
                if (e < n)
                    cp_async16(reinterpret_cast<uint4*>(&s_rec[(c % OP_RING) * OP_CHUNK + slot]) + part,
                               reinterpret_cast<const uint4*>(&recs[s_sh[e].face]) + part);
            }
            cp_async_commit();
        };
        stage(0);
        stage(1);
#ifdef B32_FILL_STATS
        if (!win) st_t1 = gtime();
#endif
        if (!mask_waited) { while (!mbar_try_wait(&s_mbar, 0)) {} mask_waited = true; }

        bool done = carry ? gdone : offscreen;             // (carry: depth-ordered windows — an early-out holds for every later window)
        const uint32_t nchunks = (n + OP_CHUNK - 1) / OP_CHUNK;
        for (uint32_t c = 0; c < nchunks; ++c) {
            cp_async_wait<1>();                               // this thread's pieces of step c have landed ...
            if (__syncthreads_and(done)) break;               // ... and so have everybody else's; slot (c+2) % 3 is free again
#ifdef B32_FILL_STATS
            if (c == 0 && !win) st_tfirst = gtime();
#endif
            stage(c + 2);
            if (done) continue;
            const SurfHot* crec = s_rec + (c % OP_RING) * OP_CHUNK;
            for (uint32_t sb = 0; sb < (uint32_t)OP_CHUNK; sb += 32) {
                const uint32_t base = c * OP_CHUNK + sb;
                if (base >= n) break;
#ifdef B32_FILL_STATS
                if (st_batches == 1) st_tb0 = gtime();       // end of batch 0
                if (st_batches == 2) st_tb1 = gtime();       // end of batch 1
                ++st_batches;
#endif
                // ---- what the weakest pixel of this block still accepts ----------------------------------------
                // painter's: wkey = smallest winner KEY in the block (0 while some pixel has no winner); z-buffer: wz =
                // largest depth in the block.  One warp reduction (REDUX) each; both halves hold the same merged state.
                uint32_t wkey = 0;
                float wz = 0.0f;
                if (!p.use_zbuffer) {
                    bool allw = __all_sync(0xFFFFFFFFu, best != 0);                          // invalid lanes carry ~0
                    wkey = allw ? __reduce_min_sync(0xFFFFFFFFu, (uint32_t)((best - 1) >> 32)) : 0u;
                } else {
                    // order-preserving image of the depth (NaN never occurs in px.z: only `z < px.z` winners are stored)
                    uint32_t zb = __float_as_uint(valid ? px.z : -INFINITY);
                    zb ^= (zb >> 31) ? 0xFFFFFFFFu : 0x80000000u;
                    zb = __reduce_max_sync(0xFFFFFFFFu, zb);
                    zb ^= (zb >> 31) ? 0x80000000u : 0xFFFFFFFFu;
                    wz = __uint_as_float(zb);
                }
                // ---- 4. early out: entries are in descending key-bucket order ------------------------------------
                // `open` pixels are those some entry of this batch or a later one could still change; only their
                // bounding box [ox0,ox1) x [oy0,oy1) needs to be met by a surface's bbox.
                uint32_t ox0 = bx0, ox1 = bx0 + OP_BW, oy0 = by0, oy1 = by0 + OP_BH;
                if (!single_batch) {
                    uint32_t k0 = s_sh[base].key;
                    if (k0 != 0xFFFFFFFFu) {
                        // upper bound of every key still to come = top of k0's bucket
                        uint64_t q0 = (k0 - kmin) >> shift;
                        if (q0 >= (uint64_t)(OP_BUCKETS - 2)) q0 = OP_BUCKETS - 1;          // bucket 1 holds the two highest quanta
                        uint64_t ub64 = (uint64_t)kmin + ((q0 + 1) << shift) - 1;
                        uint32_t ub = ub64 > 0xFFFFFFFEull ? 0xFFFFFFFEu : (uint32_t)ub64;
                        if (!p.use_zbuffer) { if (ub < wkey) { done = true; break; } }    // every later surface was drawn before every winner
                        else if (__uint_as_float(~ub) > wz) { done = true; break; }   // every later surface is behind every pixel
                        bool open = valid && (!p.use_zbuffer ? (best == 0 || (uint32_t)((best - 1) >> 32) <= ub)
                                                             : !(__uint_as_float(~ub) > px.z));
                        uint32_t om = __ballot_sync(0xFFFFFFFFu, open);                // bit q = pixel q (row-major in the block) is open
                        if (OP_DUAL) om &= 0xFFFFu;
                        if (om == 0) { done = true; break; }
                        constexpr uint32_t RM = (1u << OP_BW) - 1;                      // one block row of `om`
                        uint32_t cols = (om | (om >> OP_BW) | (om >> (2 * OP_BW)) | (om >> (3 * OP_BW))) & RM;
                        uint32_t rows = ((om & RM) ? 1u : 0u) | ((om & (RM << OP_BW)) ? 2u : 0u) | ((om & (RM << (2 * OP_BW))) ? 4u : 0u) |
                                        ((om & (RM << (3 * OP_BW))) ? 8u : 0u);
                        ox0 = bx0 + (__ffs(cols) - 1); ox1 = bx0 + (32 - __clz(cols));
                        oy0 = by0 + (__ffs(rows) - 1); oy1 = by0 + (32 - __clz(rows));
                    }
                }
                // ---- 3a. filter 32 entries, one per lane ------------------------------------------------------
                BinHead h{0, 0, 0, 0};
                bool cand = false;
                if (base + lane < n) {
                    h = s_sh[base + lane];
                    uint32_t min_x = h.bbox_x & 0xFFFF, max_x = h.bbox_x >> 16, min_y = h.bbox_y & 0xFFFF, max_y = h.bbox_y >> 16;
                    cand = !(max_x <= ox0 || min_x >= ox1 || max_y <= oy0 || min_y >= oy1);
                    if (!p.use_zbuffer) cand = cand && h.key >= wkey;
                    else cand = cand && (h.key == 0xFFFFFFFFu || !(__uint_as_float(~h.key) > wz));
                    if (cand && surface_misses_box(crec[sb + lane], ox0, ox1, oy0, oy1)) cand = false;   // exact: see surface_misses_box
                    if (PRE && cand && !(crec[sb + lane].flags & SF_FAST_EDGE) && stepped_surface_misses_box(crec[sb + lane], ox0, ox1, oy0, oy1)) cand = false;
                }
                uint32_t mask = __ballot_sync(0xFFFFFFFFu, cand);
                if (mask == 0) continue;
                uint32_t cnt = __popc(mask);
#ifdef B32_FILL_STATS
                st_surv += cnt;
#endif
                __syncwarp();
                if (cand) { uint32_t pos = __popc(mask & ((1u << lane) - 1)); my_surv[pos] = make_uint2(h.key, h.face); my_sidx[pos] = (uint8_t)(sb + lane); }
                __syncwarp();
                // ---- 3b. survivors, two at a time per pixel lane (dual: each half-warp takes every other one) -----------
                constexpr uint32_t NSUB = OP_DUAL ? 2 : 1;
                constexpr uint32_t JB = PRE ? 8u : 32u;                // survivors per shared-prefix round (4 rows x 8 = 32 lanes)
                for (uint32_t jb = 0; jb < cnt; jb += JB) {
                const uint32_t jend = PRE ? min(cnt, jb + JB) : cnt;
                float pw0 = 0.0f, pw1 = 0.0f;
                if (PRE) {                                              // lane -> (survivor jb + lane / 4, block row lane % 4)
                    const uint32_t sj = jb + (lane >> 2), yr = by0 + (lane & 3u);
                    if (sj < jend) {
                        const SurfHot& r = crec[my_sidx[sj]];
                        const uint32_t min_x = r.bbox_x & 0xFFFF, min_y = r.bbox_y & 0xFFFF, max_y = r.bbox_y >> 16;
                        if (!(r.flags & SF_FAST_EDGE) && yr >= min_y && yr < max_y) {
                            pw0 = r.w0s; pw1 = r.w1s;
                            edge_steps(pw0, pw1, r.b0, r.b1, yr - min_y);                       // row steps first (:1706-1712) ...
                            edge_steps(pw0, pw1, r.a0, r.a1, bx0 > min_x ? bx0 - min_x : 0u);   // ... then along the row to the block
                        }
                    }
                    __syncwarp();
                }
                for (uint32_t j0 = jb; j0 < jend; j0 += 2 * NSUB) {
                    #pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        uint32_t j = j0 + NSUB * k + sub;
                        float sw0 = 0.0f, sw1 = 0.0f;
                        if (PRE) {                                      // every lane takes part (no lane has left the loop)
                            const uint32_t src = ((min(j, jend - 1) - jb) << 2) | (pix / OP_BW);
                            sw0 = __shfl_sync(0xFFFFFFFFu, pw0, src); sw1 = __shfl_sync(0xFFFFFFFFu, pw1, src);
                        }
                        if (j >= jend) continue;
                        const SurfHot& r = crec[my_sidx[j]];
                        const uint2 kf = my_surv[j];
                        uint32_t min_x = r.bbox_x & 0xFFFF, max_x = r.bbox_x >> 16, min_y = r.bbox_y & 0xFFFF, max_y = r.bbox_y >> 16;
                        if (!(valid && x >= min_x && x < max_x && y >= min_y && y < max_y)) continue;
                        uint32_t c_face = kf.y + 1;
                        uint64_t c_prio = (((uint64_t)kf.x << 32) | kf.y) + 1;
                        if (!p.use_zbuffer && c_prio <= best) continue;               // drawn earlier than the current winner
                        float bc_x, bc_y, bc_z;
                        if (!(PRE ? inside_test_prefix<OP_BW>(r, x, y, sw0, sw1, bx0, bc_x, bc_y, bc_z) : inside_test(r, x, y, bc_x, bc_y, bc_z))) continue;
                        float inv_z = 0.0f, c_z = 0.0f;
                        if (p.use_zbuffer || !p.affine_textures) inv_z = bc_x * r.iz1 + bc_y * r.iz2 + bc_z * r.iz3;   // :1549
                        if (p.use_zbuffer) {
                            // lexicographic (z, face) minimum == sequential `z < zbuffer` in face order (:1553-1560, :1684)
                            c_z = 1.0f / inv_z;
                            if (!(c_z < px.z || (c_z == px.z && c_face < best_face))) continue;
                        }
#ifdef B32_FILL_STATS
                        ++st_inside;
#endif
                        // black-keyed textured surface: the texel decides whether this fragment writes (:1591-1607)
                        if ((r.flags & (SF_TEXTURED | SF_BLACK_TR)) == (SF_TEXTURED | SF_BLACK_TR)) {
                            uint32_t ti = texel_index(r, bc_x, bc_y, bc_z, inv_z, texd[r.flags >> 16], p);
                            if (ti == TEXEL_NONE || !((maskw[ti >> 5] >> (ti & 31)) & 1u)) continue;    // transparent key / black-keyed texel
                        }
                        if (!p.use_zbuffer) best = c_prio;
                        else { px.z = c_z; best_face = c_face; }
                        if (PRE) { win_bx = bc_x; win_by = bc_y; }
                    }
                }
                }
                if (OP_DUAL) {   // merge the two half-warps' winners for each pixel (exact: max / lexicographic min are associative)
                    uint64_t ob = __shfl_xor_sync(0xFFFFFFFFu, best, 16);
                    float oz = __shfl_xor_sync(0xFFFFFFFFu, px.z, 16);
                    uint32_t of = __shfl_xor_sync(0xFFFFFFFFu, best_face, 16);
                    // z-buffer: lexicographic (z, face+1) minimum, 0 = framebuffer content wins ties; painter's: max priority
                    bool take = p.use_zbuffer ? (oz < px.z || (oz == px.z && of < best_face)) : (ob > best);
                    if (PRE) {
                        const float obx = __shfl_xor_sync(0xFFFFFFFFu, win_bx, 16), oby = __shfl_xor_sync(0xFFFFFFFFu, win_by, 16);
                        if (take) { win_bx = obx; win_by = oby; }
                    }
                    if (take) { best = ob; px.z = oz; best_face = of; }
                }
            }
        }
        if (MULTI && carry) {                                          // every warp settled: the farther windows cannot change anything
            gdone = done;
            if (__syncthreads_and(gdone)) return false;
        }
        return true;
    };
    if (n_cand <= (uint32_t)OP_SORT_MAX) run_window(std::false_type{}, 0u);
    else for (uint32_t win = 0; run_window(std::true_type{}, win); ++win) {}
    cp_async_wait<0>();
    if (!mask_waited) { while (!mbar_try_wait(&s_mbar, 0)) {} }       // the bulk copy must land before the CTA exits
#ifdef B32_FILL_STATS
    st_tl = gtime();
#endif
    // ---- 5. shade each pixel's winner once (lanes 0..15 of the warp) ------------------------------------------
    if (valid && sub == 0) {
        uint32_t winner = p.use_zbuffer ? best_face : (best ? (uint32_t)((best - 1) & 0xFFFFFFFFu) + 1 : 0);
        if (winner) {
            const SurfRec& r = recs[winner - 1];
            float bc_x, bc_y, bc_z;
            if (PRE) { bc_x = win_bx; bc_y = win_by; bc_z = 1.0f - bc_x - bc_y; }      // the walk's own values (:1536-1538)
            else inside_test(r, x, y, bc_x, bc_y, bc_z);                   // same arithmetic as in the walk
            float inv_z = bc_x * r.iz1 + bc_y * r.iz2 + bc_z * r.iz3;      // :1549
            uint32_t o_r, o_g, o_b, o_blend; bool semi;
            bool wrote = RGB888 ? shade888(r, x, y, bc_x, bc_y, bc_z, inv_z, texd, reinterpret_cast<const uint32_t*>(texels), p, o_r, o_g, o_b, o_blend)
                                  : shade(r, x, y, bc_x, bc_y, bc_z, inv_z, texd, texels, p, o_r, o_g, o_b, semi);
            if (wrote) px.rgba = o_r | (o_g << 8) | (o_b << 16) | 0xFF000000u;    // pass 1: set_pixel_15 (:445-454) / set_pixel (:301-310)
#ifdef B32_FILL_STATS
            ++st_shaded;
#endif
        }
        if (px.rgba != px0.rgba) fb_rgba[y * p.width + x] = px.rgba;
        if (__float_as_uint(px.z) != __float_as_uint(px0.z)) fb_z[y * p.width + x] = px.z;
    }
    host_signal_done(p, st, HS_FILL);
#ifdef B32_FILL_STATS
    {
        for (int o = 16; o > 0; o >>= 1) { st_inside += __shfl_xor_sync(0xFFFFFFFFu, st_inside, o); st_shaded += __shfl_xor_sync(0xFFFFFFFFu, st_shaded, o); }
        if (lane == 0 && tile < 4096) {
            uint32_t* o = g_fill_stats + (tile * 16 + wt) * 8;
            o[0] = st_t0; o[1] = st_t1; o[2] = gtime(); o[3] = st_batches; o[4] = st_tfirst; o[5] = st_tb0; o[6] = st_tb1; o[7] = st_tl;
            if (slice) { o[5] = st_tprep; o[6] = st_twin; }        // crowded tiles: end of crowd_prepare / of the first crowd_next_window
        }
    }
#endif
}

#ifdef B32_FILL_STATS
extern "C" int b32_debug_fill_stats(uint32_t* out, uint32_t n_words) {
    return (int)cudaMemcpyFromSymbol(out, g_fill_stats, (size_t)n_words * 4);
}
#endif

// =================================================================================================
// ordered pass (pass 2 + x-ray): per-tile sort by the unique draw-order key, then strict in-order replay
// =================================================================================================
// The write stage of rasterize_triangle_15 (render.rs:1664-1702) against a pixel held in registers.
__device__ __forceinline__ void write_ordered(const SurfRec& r, Pixel& px, float z, uint32_t o_r, uint32_t o_g, uint32_t o_b, bool semi,
                                              const CallParams& p) {
    uint32_t editor_alpha = (r.flags >> 8) & 0xFF;
    if (editor_alpha == 0) return;                                                 // :1664-1669
    uint32_t br = px.rgba & 0xFF, bg = (px.rgba >> 8) & 0xFF, bb = (px.rgba >> 16) & 0xFF;
    uint32_t blend_mode = r.flags & SF_BLEND_MASK;
    if (p.xray_mode) {                                                             // :507-526
        px.rgba = ((o_r + br) >> 1) | (((o_g + bg) >> 1) << 8) | (((o_b + bb) >> 1) << 16) | 0xFF000000u;
        return;
    }
    bool skip_z_write = (r.flags & SF_TRANSPARENT) != 0;                           // pass 2, :2561-2569
    if (p.use_zbuffer) {
        if (editor_alpha < 255) { /* rejected only on z >= zbuffer (:604), already tested by the caller */ }
        else if (!(z < px.z)) return;                                              // :1684
        if (!skip_z_write) px.z = z;
    }
    if (semi && blend_mode != B32_BLEND_OPAQUE) {                                  // :1686 / :578 / :1697
        o_r = blend5(o_r, br, blend_mode); o_g = blend5(o_g, bg, blend_mode); o_b = blend5(o_b, bb, blend_mode);
    }
    if (editor_alpha < 255) {                                                      // :587-593
        uint32_t a = editor_alpha, ia = 255 - a;
        o_r = (o_r * a + br * ia) / 255; o_g = (o_g * a + bg * ia) / 255; o_b = (o_b * a + bb * ia) / 255;
    }
    px.rgba = o_r | (o_g << 8) | (o_b << 16) | 0xFF000000u;
}

// The write stage of rasterize_triangle (RGB888, render.rs:1392-1424) against a pixel held in registers:
// set_pixel / set_pixel_blended / set_pixel_with_depth / the two editor-alpha writers (render.rs:301-437).
__device__ __forceinline__ void write_ordered888(const SurfRec& r, Pixel& px, float z, uint32_t o_r, uint32_t o_g, uint32_t o_b, uint32_t blend,
                                                 const CallParams& p) {
    uint32_t editor_alpha = (r.flags >> 8) & 0xFF;
    if (editor_alpha == 0) return;                                                 // :1392-1398
    uint32_t br = px.rgba & 0xFF, bg = (px.rgba >> 8) & 0xFF, bb = (px.rgba >> 16) & 0xFF;
    if (p.use_zbuffer) {
        if (editor_alpha < 255) { if (z >= px.z) return; }                         // :393
        else if (!(z < px.z)) return;                                              // :425, :1408
        px.z = z;
    }
    if (blend != B32_BLEND_OPAQUE) { o_r = blend8(o_r, br, blend); o_g = blend8(o_g, bg, blend); o_b = blend8(o_b, bb, blend); }
    if (editor_alpha < 255) {                                                      // :362-371: float lerp, `as u8`
        float a = (float)editor_alpha / 255.0f;
        float inv_a = 1.0f - a;
        o_r = f2u8((float)o_r * a + (float)br * inv_a);
        o_g = f2u8((float)o_g * a + (float)bg * inv_a);
        o_b = f2u8((float)o_b * a + (float)bb * inv_a);
    }
    px.rgba = o_r | (o_g << 8) | (o_b << 16) | 0xFF000000u;
}

#ifndef B32_ORD_SUB
#define B32_ORD_SUB 2
#endif
constexpr int ORD_SUB = B32_ORD_SUB; // a step (one CTA barrier) = ORD_SUB batches of 32 surfaces: fewer barriers, and the warps' loads even out over a step
constexpr int ORD_CHUNK = 32 * ORD_SUB;   // surfaces staged in shared memory per step (128 B each), ORD_SUB 16-byte pieces per thread
constexpr int ORD_RING = 3;          // steps c, c+1, c+2 in flight
#ifndef B32_ORD_GROUP
#define B32_ORD_GROUP 2
#endif
constexpr int ORD_GROUP = B32_ORD_GROUP;   // fragments of one pixel whose texels are requested together
constexpr size_t ORD_SMEM = (size_t)ORD_SORT_MAX * sizeof(BinHead) + (size_t)ORD_RING * ORD_CHUNK * sizeof(SurfRec) + (FILL_THREADS / 32) * 32;
// PRE instantiations (shared edge prefix, see k_fill_opaque): the edge values of every (survivor, block row) of a batch, per warp
constexpr size_t ORD_SMEM_PRE = ORD_SMEM + (size_t)(FILL_THREADS / 32) * 32 * 4 * sizeof(float2);
static_assert(ORD_SMEM % 8 == 0, "the prefix table behind it holds float2");
static_assert(FILL_THREADS == 32 * 8, "one 16-byte piece of a batch's staged records per thread");

__device__ __forceinline__ uint64_t ord_key(const BinHead& h) { return ((uint64_t)h.key << 32) | h.face; }

// CTA-wide bitonic sort (ascending by draw-order key) of m = 2^k entries; works on shared or global memory.
__device__ void bitonic_sort_heads(BinHead* a, uint32_t m) {
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                uint32_t l = i ^ j;
                if (l > i) {
                    BinHead x = a[i], y = a[l];
                    bool asc = (i & k) == 0;
                    if ((ord_key(x) > ord_key(y)) == asc) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
}

// One CTA per 16x16 tile, one warp per 8x4 block, one lane per pixel.  The tile's bin is sorted by the unique
// draw-order key; the records stream through a 3-deep cp.async ring in that order; every pixel replays the surfaces
// covering it one after the other, applying the reference's z-test / blend / write rules (render.rs:1664-1702 and, for
// RGB888, :1392-1424) to a colour + depth held in registers.  Per step each lane first filters one of the 32 staged
// surfaces against the warp's block (bbox, exact corner trivial reject); every pixel then marks which survivors cover
// it (one bit each) and folds its own fragments in draw order — the order only matters per pixel, so the lanes of a
// warp shade different surfaces side by side.
// PRE (float / ortho calls, or the stepped-surface hint): surfaces whose edge values are the reference's rounded additions
// get their chains replayed once per (survivor, block row) — 32 chains side by side, into a per-warp table — instead of once
// per covered fragment per pixel; phase A then tests them exactly (not just by bounding box) and phase B adds < 8 steps.
template <bool RGB888, bool PRE>
__global__ void __launch_bounds__(FILL_THREADS)
k_fill_ordered(const SurfRec* __restrict__ recs, const uint4* __restrict__ masks, BinHead* __restrict__ scratch,
               const uint64_t* __restrict__ keys, const TexDev* __restrict__ tex, const void* __restrict__ texels,
               uint32_t* __restrict__ fb_rgba, float* __restrict__ fb_z, CallState* __restrict__ st, uint32_t* __restrict__ sticky,
               CallParams p, uint32_t scratch_cap) {
    extern __shared__ __align__(128) uint8_t ord_smem[];
    SurfRec* s_rec = reinterpret_cast<SurfRec*>(ord_smem);                              // [ORD_RING][ORD_CHUNK]
    BinHead* s_sorted = reinterpret_cast<BinHead*>(s_rec + ORD_RING * ORD_CHUNK);       // [ORD_SORT_MAX]
    uint8_t* s_sidx = reinterpret_cast<uint8_t*>(s_sorted + ORD_SORT_MAX);              // [warps][32] survivors of the step, in order
    float2* s_pre = reinterpret_cast<float2*>(ord_smem + ORD_SMEM);                     // PRE: [warps][32 survivors][4 block rows] (w0, w1) at the row's first column
    uint32_t* s_cand = reinterpret_cast<uint32_t*>(s_rec);                              // [ORD_SORT_MAX] face indices of a window (the ring is idle then)
    __shared__ uint32_t s_wsum[FILL_THREADS / 32];
    __shared__ uint32_t s_n;
    if (p.enq_ordered) pdl_wait();                      // enqueued right behind pass 1: k_fill_opaque (and k_setup before it) have completed
    host_stamp_start(p, HS_ORDERED);
    const bool all_ordered = RGB888 ? true : p.xray_mode != 0;      // every drawn surface is replayed (RGB888: this kernel only runs when some surface may blend)
    {
        CallState s = *st;
        bool nothing = call_aborts(s, p.use_zbuffer, RGB888);
        // enqueue-only callers learn of the reference's panics at the next sync; in x-ray mode this is the frame's only fill kernel
        if (nothing && p.async_call && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(sticky, s.oob == 2 ? 32u : s.oob ? 1u : 2u);
        const uint32_t n_ordered = all_ordered ? s.n_opaque + s.n_transp : s.n_transp;
        nothing = nothing || n_ordered == 0 || (RGB888 && s.n_transp == 0);             // nothing to replay (the usual case of an enqueued frame)
        // k_setup counted the ordered entries per mask tile: a tile with more than fit shared memory needs a slice of the
        // global scratch; if that is too small NO tile draws (a blocking call sizes it before it launches this kernel;
        // enqueue-only callers get an error)
        if (!nothing && s.obin_max > (uint32_t)ORD_SORT_MAX && s.obin_max > scratch_cap) {
            if (blockIdx.x == 0 && threadIdx.x == 0) { st->obin_overflow = 1; if (p.async_call) atomicOr(sticky, 16u); }
            nothing = true;
        }
        if (nothing) { host_signal_done(p, st, HS_ORDERED); return; }
    }
    const uint32_t tile = blockIdx.x;
    const uint32_t ttx = tile % p.tiles_x, tty = tile / p.tiles_x;
    const uint32_t mtile = (tty >> p.mshift) * p.mtiles_x + (ttx >> p.mshift);
    const uint4* mrow = masks + (size_t)mtile * p.n_groups;
    CandScan cs = cand_scan<FILL_THREADS>(mrow, p.n_groups, s_wsum);
    if (cs.total == 0) { host_signal_done(p, st, HS_ORDERED); return; }
    // ---- this tile's draw-order entries: the candidates that are in the ordered pass and touch the tile.  The unique
    //      64-bit key (pass:2 | depth key:32 | face:30) = opaque list first (sorted only in painter's mode), then the
    //      transparent list, ties by face index = stable sort (render.rs:2522-2542); RGB888: one list, no pass bit.
    // The entries are collected in shared memory AND (when the tile has a slice of the scratch) in global memory: how many
    // of the tile's candidates are ordered entries is only known afterwards; k_setup's count per mask tile (obin_max, checked
    // above) guarantees a slice of at least that many whenever more than ORD_SORT_MAX can occur.
    BinHead* gslice = scratch + (size_t)tile * scratch_cap;
    if (threadIdx.x == 0) s_n = 0;
    for (uint32_t w0 = 0; w0 < cs.total; w0 += ORD_SORT_MAX) {
        const uint32_t wn = min((uint32_t)ORD_SORT_MAX, cs.total - w0);
        cand_expand(mrow, cs, w0, wn, s_cand);           // (its barrier also publishes s_n = 0)
        for (uint32_t i = threadIdx.x; i < wn; i += blockDim.x) {
            const uint32_t fi = s_cand[i];
            const uint64_t k64 = keys[fi];
            const uint32_t cls = (uint32_t)(k64 >> 32);
            if (!(cls < 2 && (cls == 1 || all_ordered))) continue;
            const uint2 bb = *reinterpret_cast<const uint2*>(&recs[fi].bbox_x);               // all zero = empty surface
            if (!bb.x) continue;
            BinHead h{bb.x, bb.y, 0, 0};
            uint32_t tx0, tx1, ty0, ty1;
            bbox_mtiles(bb.x, bb.y, 4u, tx0, tx1, ty0, ty1);
            if (!(ttx >= tx0 && ttx <= tx1 && tty >= ty0 && tty <= ty1)) continue;
            const uint64_t okey = ((uint64_t)(RGB888 ? 0u : cls) << 62) | ((uint64_t)(uint32_t)k64 << 30) | fi;
            h.key = (uint32_t)(okey >> 32); h.face = (uint32_t)okey;
            const uint32_t pos = atomicAdd(&s_n, 1u);
            if (pos < (uint32_t)ORD_SORT_MAX) s_sorted[pos] = h;
            if (pos < scratch_cap) gslice[pos] = h;
        }
        __syncthreads();
    }
    const uint32_t n = s_n;
    if (n == 0) { host_signal_done(p, st, HS_ORDERED); return; }
    uint32_t m = 2;
    while (m < n) m <<= 1;                              // scratch_cap is a power of two >= n whenever the scratch is used
    BinHead* sorted;
    if (m <= (uint32_t)ORD_SORT_MAX) {
        for (uint32_t i = n + threadIdx.x; i < m; i += blockDim.x) s_sorted[i] = BinHead{0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu};
        sorted = s_sorted;
    } else {
        for (uint32_t i = n + threadIdx.x; i < m; i += blockDim.x) gslice[i] = BinHead{0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu};
        sorted = gslice;
    }
    __syncthreads();
    bitonic_sort_heads(sorted, m);

    const uint32_t tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bx0 = tx * TILE_W + (warp & 1) * 8, by0 = ty * TILE_H + (warp >> 1) * 4;
    const uint32_t x = bx0 + (lane & 7), y = by0 + (lane >> 3);
    const bool valid = x < p.width && y < p.height;
    Pixel px{0, 0.0f};
    if (valid) { px.rgba = fb_rgba[y * p.width + x]; px.z = fb_z[y * p.width + x]; }
    const Pixel px0 = px;
    uint8_t* my_sidx = s_sidx + warp * 32;
    const bool early_z = p.use_zbuffer && !p.xray_mode;                                // :1553-1560 / :1313-1320

    auto stage = [&](uint32_t c) {                                                     // step c -> ring slot c % ORD_RING
        #pragma unroll
        for (int sb = 0; sb < ORD_SUB; ++sb) {
            uint32_t slot = sb * 32 + (threadIdx.x >> 3), e = c * ORD_CHUNK + slot;
            if (e < n) {
                uint32_t f = sorted[e].face & 0x3FFFFFFFu;
                cp_async16(reinterpret_cast<uint4*>(&s_rec[(c % ORD_RING) * ORD_CHUNK + slot]) + (threadIdx.x & 7),
                           reinterpret_cast<const uint4*>(&recs[f]) + (threadIdx.x & 7));
            }
        }
        cp_async_commit();
    };
    stage(0);
    stage(1);
    const uint32_t nchunks = (n + ORD_CHUNK - 1) / ORD_CHUNK;
    for (uint32_t c = 0; c < nchunks; ++c) {
        cp_async_wait<1>();
        __syncthreads();                                   // step c has landed for everybody; slot (c+2) % 3 is free again
        stage(c + 2);
        for (uint32_t sb = 0; sb < (uint32_t)ORD_SUB && c * ORD_CHUNK + sb * 32 < n; ++sb) {
        const SurfRec* crec = s_rec + (c % ORD_RING) * ORD_CHUNK + sb * 32;
        // ---- filter: lane = one staged surface vs this warp's 8x4 block ----
        bool cand = false;
        if (c * ORD_CHUNK + sb * 32 + lane < n && bx0 < p.width && by0 < p.height) {
            const SurfRec& r = crec[lane];
            uint32_t min_x = r.bbox_x & 0xFFFF, max_x = r.bbox_x >> 16, min_y = r.bbox_y & 0xFFFF, max_y = r.bbox_y >> 16;
            cand = !(max_x <= bx0 || min_x >= bx0 + 8 || max_y <= by0 || min_y >= by0 + 4);
            if (cand && surface_misses_box(r, bx0, bx0 + 8, by0, by0 + 4)) cand = false;
        }
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, cand);
        if (mask == 0) continue;
        const uint32_t cnt = __popc(mask);
        __syncwarp();
        if (cand) my_sidx[__popc(mask & ((1u << lane) - 1))] = (uint8_t)lane;          // ascending = draw order
        __syncwarp();
        // ---- A. which of the step's survivors cover this pixel: the inside tests, in lockstep (nothing here depends on the
        //         pixel's running colour / depth).  Surfaces on the replayed-additions edge path are only bbox-tested here.
        float2* my_pre = s_pre + (size_t)(threadIdx.x >> 5) * 32 * 4;
        if (PRE) {                                             // lane -> (survivor jb + lane / 4, block row lane % 4)
            for (uint32_t jb = 0; jb < cnt; jb += 8) {
                const uint32_t sj = jb + (lane >> 2), yr = by0 + (lane & 3u);
                if (sj < cnt) {
                    const SurfRec& r = crec[my_sidx[sj]];
                    const uint32_t min_x = r.bbox_x & 0xFFFF, min_y = r.bbox_y & 0xFFFF, max_y = r.bbox_y >> 16;
                    if (!(r.flags & SF_FAST_EDGE) && yr >= min_y && yr < max_y) {
                        float w0 = r.w0s, w1 = r.w1s;
                        edge_steps(w0, w1, r.b0, r.b1, yr - min_y);                        // row steps first (:1706-1712) ...
                        edge_steps(w0, w1, r.a0, r.a1, bx0 > min_x ? bx0 - min_x : 0u);    // ... then along the row to the block
                        my_pre[sj * 4 + (lane & 3u)] = make_float2(w0, w1);
                    }
                }
            }
            __syncwarp();
        }
        uint32_t cov = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
            const SurfRec& r = crec[my_sidx[j]];
            uint32_t min_x = r.bbox_x & 0xFFFF, max_x = r.bbox_x >> 16, min_y = r.bbox_y & 0xFFFF, max_y = r.bbox_y >> 16;
            if (!(valid && x >= min_x && x < max_x && y >= min_y && y < max_y)) continue;
            float bc_x, bc_y, bc_z;
            if (PRE) {
                const float2 pw = my_pre[j * 4 + (y - by0)];
                if (inside_test_prefix<8>(r, x, y, pw.x, pw.y, bx0, bc_x, bc_y, bc_z)) cov |= 1u << j;
            } else if (!(r.flags & SF_FAST_EDGE) || inside_test(r, x, y, bc_x, bc_y, bc_z)) cov |= 1u << j;
        }
        // ---- B. every pixel folds ITS fragments in draw order, ORD_GROUP at a time: lanes work on different surfaces in
        //         the same instruction (a small triangle covers a quarter of the block: walking the survivors in lockstep
        //         left three lanes in four idle through the shading).  The inside tests, depths and texel requests of a
        //         group are issued together; only then are its fragments shaded and written one after the other.
        while (cov) {
            bool in[ORD_GROUP]; float fbx[ORD_GROUP], fby[ORD_GROUP], fz[ORD_GROUP]; uint32_t ftex[ORD_GROUP], fidx[ORD_GROUP];
            bool nan_before = false;                       // an earlier fragment of this group has a NaN depth (see below)
            #pragma unroll
            for (int k = 0; k < ORD_GROUP; ++k) {
                in[k] = false; fbx[k] = fby[k] = fz[k] = 0.0f; ftex[k] = 0; fidx[k] = 0;
                if (!cov) continue;
                const uint32_t j = __ffs(cov) - 1; cov &= cov - 1;
                fidx[k] = my_sidx[j];
                const SurfRec& r = crec[fidx[k]];
                float bc_x, bc_y, bc_z;
                if (PRE) {
                    const float2 pw = my_pre[j * 4 + (y - by0)];
                    if (!inside_test_prefix<8>(r, x, y, pw.x, pw.y, bx0, bc_x, bc_y, bc_z)) continue;      // (cannot fail: phase A tested it)
                } else if (!inside_test(r, x, y, bc_x, bc_y, bc_z)) continue;           // (only the slow edge path can still fail)
                float inv_z = bc_x * r.iz1 + bc_y * r.iz2 + bc_z * r.iz3;              // :1549
                float z = 1.0f / inv_z;
                // A reject against the depth the pixel has NOW is a reject at the fragment's own time as long as the depth only
                // decreases until then.  The one way it can rise is through a NaN: render_mesh's editor-alpha writer stores a
                // NaN depth (`z >= zbuffer` is false, render.rs:393) and the next such fragment replaces it with any finite
                // one.  A NaN depth can only come from a fragment whose own depth is NaN: after one of those in this group
                // nothing is rejected early (the fold below applies the reference's test at its time).
                if (early_z && !nan_before && z >= px.z) continue;
                nan_before = nan_before || z != z;
                in[k] = true; fbx[k] = bc_x; fby[k] = bc_y; fz[k] = z;
                ftex[k] = sample_texel<RGB888>(r, bc_x, bc_y, bc_z, inv_z, tex, texels, p);
            }
            #pragma unroll
            for (int k = 0; k < ORD_GROUP; ++k) {          // the fold over the fragments, in draw order
                if (!in[k]) continue;
                if (early_z && fz[k] >= px.z) continue;                                // the reference's early test, at its time
                const SurfRec& r = crec[fidx[k]];
                const float bc_x = fbx[k], bc_y = fby[k], bc_z = 1.0f - bc_x - bc_y;   // same expression as inside_test
                uint32_t o_r, o_g, o_b;
                if (RGB888) {
                    uint32_t blend;
                    if (!shade_color888(r, x, y, bc_x, bc_y, bc_z, ftex[k], p, o_r, o_g, o_b, blend)) continue;
                    write_ordered888(r, px, fz[k], o_r, o_g, o_b, blend, p);
                } else {
                    bool semi;
                    if (!shade_color(r, x, y, bc_x, bc_y, bc_z, ftex[k], p, o_r, o_g, o_b, semi)) continue;
                    write_ordered(r, px, fz[k], o_r, o_g, o_b, semi, p);
                }
            }
        }
        }   // batch
    }
    cp_async_wait<0>();
    if (valid) {
        if (px.rgba != px0.rgba) fb_rgba[y * p.width + x] = px.rgba;
        if (__float_as_uint(px.z) != __float_as_uint(px0.z)) fb_z[y * p.width + x] = px.z;
    }
    host_signal_done(p, st, HS_ORDERED);
}

// =================================================================================================
// wireframe phase (render.rs:2574-2635): editor feature, off in RasterSettings::game()
// =================================================================================================
// Every line of one wireframe pass has the same colour and only READS the z-buffer, so the order of the
// lines does not matter and each unique edge can be walked by its own thread.  What does matter is the
// reference's de-duplication: edges are compared by their integer end points only and the FIRST
// occurrence (in face order) keeps its depths (:2589-2591), so edge e is drawn iff no earlier edge of
// the same kind has the same end points.
struct WireEdge { int32_t x0, y0, x1, y1; float z0, z1; };

__device__ __forceinline__ bool wire_edge(const WireTri& t, uint32_t k, WireEdge& e) {
    uint32_t a = k, b = (k + 1) % 3;
    int32_t ax = f2i32(t.x[a]), ay = f2i32(t.y[a]), bx = f2i32(t.x[b]), by = f2i32(t.y[b]);      // `as i32` (:2581-2585)
    bool lt = ax < bx || (ax == bx && ay < by);                                                   // (x0,y0) < (x1,y1)
    e = lt ? WireEdge{ax, ay, bx, by, t.z[a], t.z[b]} : WireEdge{bx, by, ax, ay, t.z[b], t.z[a]};
    return true;
}

// The reference keeps the FIRST edge (in face order) of every distinct integer end-point pair with a linear search per
// edge (`unique_edges.iter().any(..)`, :2589: O(n^2)).  Same set, O(n): an open-addressing table in which every slot only
// ever holds edges of ONE end-point pair and converges to the smallest edge index of that pair (atomicMin); edge e is
// drawn iff its pair's slot holds e.  A slot is claimed by compare-and-swap from EMPTY; a later edge either finds its own
// pair there (atomicMin) or a different one (probe on); nothing is ever removed, so all edges of a pair meet in one slot.
constexpr uint32_t WIRE_EMPTY = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t wire_hash(const WireEdge& e) {
    uint32_t h = (uint32_t)e.x0 * 0x9E3779B1u ^ (uint32_t)e.y0 * 0x85EBCA77u ^ (uint32_t)e.x1 * 0xC2B2AE3Du ^ (uint32_t)e.y1 * 0x27D4EB2Fu;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}
__device__ __forceinline__ bool wire_same(const WireEdge& a, const WireEdge& b) { return a.x0 == b.x0 && a.y0 == b.y0 && a.x1 == b.x1 && a.y1 == b.y1; }

__global__ void __launch_bounds__(128)
k_wire_dedup(const WireTri* __restrict__ wire, uint32_t nf, uint32_t kind, uint32_t* __restrict__ table, uint32_t mask) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < nf * 3; e += gridDim.x * blockDim.x) {
        const WireTri& t = wire[e / 3];
        if (t.kind != kind) continue;
        WireEdge me;
        wire_edge(t, e % 3, me);
        for (uint32_t h = wire_hash(me) & mask;; h = (h + 1) & mask) {
            uint32_t cur = table[h];
            if (cur == WIRE_EMPTY) {
                cur = atomicCAS(&table[h], WIRE_EMPTY, e);
                if (cur == WIRE_EMPTY) break;                               // claimed for this end-point pair
            }
            WireEdge oe;
            wire_edge(wire[cur / 3], cur % 3, oe);                          // any edge ever stored here has the slot's pair
            if (wire_same(oe, me)) { atomicMin(&table[h], e); break; }
        }
    }
}

// draw_line_3d (depth_test = true, :768-817) / draw_line (:714-751); colour via set_pixel (:301-310)
__global__ void __launch_bounds__(128)
k_wire(const WireTri* __restrict__ wire, uint32_t nf, uint32_t kind, uint32_t color, bool depth_test,
       const uint32_t* __restrict__ table, uint32_t mask,
       uint32_t* __restrict__ fb_rgba, const float* __restrict__ fb_z, CallState* __restrict__ st, CallParams p) {
    if (call_aborts(*st, p.use_zbuffer, p.rgb888)) return;
    const int32_t W = (int32_t)p.width, H = (int32_t)p.height;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < nf * 3; e += gridDim.x * blockDim.x) {
        const WireTri& t = wire[e / 3];
        if (t.kind != kind) continue;
        WireEdge me;
        wire_edge(t, e % 3, me);
        uint32_t first = WIRE_EMPTY;                                        // first occurrence of this end-point pair (:2589)
        for (uint32_t h = wire_hash(me) & mask;; h = (h + 1) & mask) {
            uint32_t cur = table[h];
            if (cur == WIRE_EMPTY) break;                                   // cannot happen (k_wire_dedup inserted every edge): draw nothing
            WireEdge oe;
            wire_edge(wire[cur / 3], cur % 3, oe);
            if (wire_same(oe, me)) { first = cur; break; }
        }
        if (first != e) continue;
        // Bresenham with wrapping i32 arithmetic (release-mode Rust)
        int32_t x0 = me.x0, y0 = me.y0, x1 = me.x1, y1 = me.y1;
        int32_t ddx = (int32_t)((uint32_t)x1 - (uint32_t)x0), ddy = (int32_t)((uint32_t)y1 - (uint32_t)y0);
        int32_t dx = ddx < 0 ? (int32_t)(0u - (uint32_t)ddx) : ddx;
        int32_t dy = -(ddy < 0 ? (int32_t)(0u - (uint32_t)ddy) : ddy);
        int32_t sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
        int32_t err = (int32_t)((uint32_t)dx + (uint32_t)dy), x = x0, y = y0;
        float total_steps = (float)max(dx, max(-dy, 1));
        float step = 0.0f;
        // The reference walks every step of the line, on screen or not: end points saturated from non-finite or absurdly
        // large coordinates mean billions of iterations (seconds of CPU there, a hung SM here).  Such an edge is not
        // walked; the call reports B32_ERR_UNSUPPORTED.  (4.12 fixed-point projection cannot exceed 2^20 steps.)
        if ((uint32_t)max(dx, -dy) > WIRE_MAX_STEPS || dx < 0 || dy > 0) { st->wire_too_long = 1; continue; }
        for (;;) {
            if (x >= 0 && x < W && y >= 0 && y < H) {
                uint32_t idx = (uint32_t)y * p.width + (uint32_t)x;
                bool pass = true;
                if (depth_test) {
                    float tt = step / total_steps;
                    float z = me.z0 + tt * (me.z1 - me.z0);
                    pass = z < fb_z[idx];
                }
                if (pass) fb_rgba[idx] = color;
            }
            if (x == x1 && y == y1) break;
            int32_t e2 = (int32_t)(2u * (uint32_t)err);
            if (e2 >= dy) { err = (int32_t)((uint32_t)err + (uint32_t)dy); x += sx; if (depth_test) step += 1.0f; }
            if (e2 <= dx) { err = (int32_t)((uint32_t)err + (uint32_t)dx); y += sy; if (depth_test && e2 < dy) step += 1.0f; }
        }
    }
}

// =================================================================================================
// skybox sphere pass: Framebuffer::render_skybox step 1 (render.rs:89-139) + rasterize_skybox_triangle (:242-299)
// =================================================================================================
// Faces are drawn in order into pixels nothing else reads (no depth, no blending), so a pixel's final colour is that
// of the LAST face covering its centre: order-free like pass 1, with the face index as priority.  k_sky_setup
// projects + culls one face per thread and emits a bin head (walked into the tile bins by k_bin_opaque);
// k_sky_fill gives each pixel of a tile to one thread.
__global__ void __launch_bounds__(SETUP_GROUP)
k_sky_setup(const b32_sky_vertex* __restrict__ verts, const uint32_t* __restrict__ faces, SkyRec* __restrict__ recs,
            BinHead* __restrict__ heads, uint4* __restrict__ masks, CallState* __restrict__ st, uint32_t* __restrict__ zero_next, uint32_t zero_words, CallParams p) {
    extern __shared__ __align__(16) uint8_t su_smem[];
    uint4* s_mask = reinterpret_cast<uint4*>(su_smem);
    const uint32_t n_mtiles = p.mtiles_x * p.mtiles_y;
    pdl_launch_dependents();
    if (blockIdx.x == 0) for (uint32_t i = threadIdx.x; i < zero_words; i += blockDim.x) zero_next[i] = 0;
    for (uint32_t group = blockIdx.x; group < p.n_groups; group += gridDim.x) {
        const uint32_t fi = group * SETUP_GROUP + threadIdx.x;
        masks_zero(s_mask, n_mtiles);
        __syncthreads();
        BinHead head{0, 0, 0, fi};
        if (fi < p.nf) do {
            uint32_t i0 = faces[fi * 3], i1 = faces[fi * 3 + 1], i2 = faces[fi * 3 + 2];
            if (i0 >= p.nv || i1 >= p.nv || i2 >= p.nv) { st->oob = 1; break; }
            const b32_sky_vertex a = verts[i0], b = verts[i1], c = verts[i2];
            // rel_pos, perspective_transform, `cam_space.z <= 0.1` => NaN marker, project (:96-108): the float path
            TVert t0 = transform_vertex(a.pos[0], a.pos[1], a.pos[2], p, nullptr);
            TVert t1 = transform_vertex(b.pos[0], b.pos[1], b.pos[2], p, nullptr);
            TVert t2 = transform_vertex(c.pos[0], c.pos[1], c.pos[2], p, nullptr);
            if (t0.w <= 0.1f || t1.w <= 0.1f || t2.w <= 0.1f) break;                              // :118-120
            if (t0.x != t0.x || t1.x != t1.x || t2.x != t2.x) break;                              // a NaN x also drops the face
            float signed_area = (t1.x - t0.x) * (t2.y - t0.y) - (t2.x - t0.x) * (t1.y - t0.y);    // :124
            if (signed_area >= 0.0f) break;
            // rasterize_skybox_triangle: inclusive bbox (:252-259), degenerate test (:262-266)
            uint32_t min_x = f2u32sat(fmaxf(fminf(fminf(t0.x, t1.x), t2.x), 0.0f));
            uint32_t max_x = f2u32sat(fminf(fmaxf(fmaxf(t0.x, t1.x), t2.x), (float)p.width - 1.0f));
            uint32_t min_y = f2u32sat(fmaxf(fminf(fminf(t0.y, t1.y), t2.y), 0.0f));
            uint32_t max_y = f2u32sat(fminf(fmaxf(fmaxf(t0.y, t1.y), t2.y), (float)p.height - 1.0f));
            if (min_x > max_x || min_y > max_y) break;
            if (max_x >= p.width || max_y >= p.height) break;      // only for NaN-free garbage (w or h = 0 never reaches here)
            float denom = (t1.y - t2.y) * (t0.x - t2.x) + (t2.x - t1.x) * (t0.y - t2.y);
            if (fabsf(denom) < 0.0001f) break;
            SkyRec r;
            r.p0x = t0.x; r.p0y = t0.y; r.p1x = t1.x; r.p1y = t1.y; r.p2x = t2.x; r.p2y = t2.y;
            r.inv_denom = 1.0f / denom;
            r.c0 = a.r | (a.g << 8) | (a.b << 16); r.c1 = b.r | (b.g << 8) | (b.b << 16); r.c2 = c.r | (c.g << 8) | (c.b << 16);
            r._pad0 = r._pad1 = 0;
            recs[fi] = r;
            head = BinHead{min_x | ((max_x + 1) << 16), min_y | ((max_y + 1) << 16), 0u, fi};   // exclusive max, as the mesh heads
        } while (0);
        if (fi < p.nf) heads[fi] = head;
        masks_mark(s_mask, head.bbox_x, head.bbox_y, head.bbox_x != 0, p);
        __syncthreads();
        masks_flush(s_mask, masks, n_mtiles, group, p, make_uint4(0, 0, 0, 0), nullptr, st);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(FILL_THREADS)
k_sky_fill(const SkyRec* __restrict__ recs, const uint4* __restrict__ masks, const BinHead* __restrict__ heads,
           uint32_t* __restrict__ fb_rgba, const CallState* __restrict__ st, CallParams p) {
    __shared__ BinHead s_head[FILL_THREADS];
    __shared__ SkyRec s_rec[FILL_THREADS];
    __shared__ uint32_t s_cand[FILL_THREADS];
    __shared__ uint32_t s_wsum[FILL_THREADS / 32];
    pdl_wait();
    if (st->oob) return;
    const uint32_t tile = blockIdx.x;
    const uint32_t ttx = tile % p.tiles_x, tty = tile / p.tiles_x;
    const uint4* mrow = masks + (size_t)((tty >> p.mshift) * p.mtiles_x + (ttx >> p.mshift)) * p.n_groups;
    CandScan cs = cand_scan<FILL_THREADS>(mrow, p.n_groups, s_wsum);
    if (cs.total == 0) return;
    const uint32_t x = ttx * TILE_W + (threadIdx.x % TILE_W), y = tty * TILE_H + (threadIdx.x / TILE_W);
    const bool valid = x < p.width && y < p.height;
    const float px = (float)x + 0.5f, py = (float)y + 0.5f;                                      // :270-271
    uint32_t best = 0;                                       // face + 1 of the last face covering this pixel centre
    // The candidates are walked FILL_THREADS at a time; a face below the best one so far is skipped by its index alone.
    // Heads and records of a step sit in shared memory.
    for (uint32_t rem = cs.total; rem > 0;) {
        const uint32_t cnt = min((uint32_t)FILL_THREADS, rem);
        rem -= cnt;
        __syncthreads();
        cand_expand(mrow, cs, rem, cnt, s_cand);
        if (threadIdx.x < cnt) { const BinHead h = heads[s_cand[threadIdx.x]]; s_head[threadIdx.x] = h; s_rec[threadIdx.x] = recs[h.face]; }
        __syncthreads();
        for (uint32_t i = cnt; i-- > 0;) {
            const BinHead h = s_head[i];
            if (h.face + 1 <= best) continue;
            if (!(valid && x >= (h.bbox_x & 0xFFFF) && x < (h.bbox_x >> 16) && y >= (h.bbox_y & 0xFFFF) && y < (h.bbox_y >> 16))) continue;
            const SkyRec& r = s_rec[i];
            float w0 = ((r.p1y - r.p2y) * (px - r.p2x) + (r.p2x - r.p1x) * (py - r.p2y)) * r.inv_denom;   // :274-276
            float w1 = ((r.p2y - r.p0y) * (px - r.p2x) + (r.p0x - r.p2x) * (py - r.p2y)) * r.inv_denom;
            float w2 = 1.0f - w0 - w1;
            if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) best = h.face + 1;
        }
    }
    if (best) {
        const SkyRec r = recs[best - 1];
        float w0 = ((r.p1y - r.p2y) * (px - r.p2x) + (r.p2x - r.p1x) * (py - r.p2y)) * r.inv_denom;
        float w1 = ((r.p2y - r.p0y) * (px - r.p2x) + (r.p0x - r.p2x) * (py - r.p2y)) * r.inv_denom;
        float w2 = 1.0f - w0 - w1;
        uint32_t cr = f2u8((float)(r.c0 & 0xFF) * w0 + (float)(r.c1 & 0xFF) * w1 + (float)(r.c2 & 0xFF) * w2);          // :281-283
        uint32_t cg = f2u8((float)((r.c0 >> 8) & 0xFF) * w0 + (float)((r.c1 >> 8) & 0xFF) * w1 + (float)((r.c2 >> 8) & 0xFF) * w2);
        uint32_t cb = f2u8((float)(r.c0 >> 16) * w0 + (float)(r.c1 >> 16) * w1 + (float)(r.c2 >> 16) * w2);
        fb_rgba[y * p.width + x] = cr | (cg << 8) | (cb << 16) | 0xFF000000u;
    }
}

// =================================================================================================
// star field: Framebuffer::render_skybox step 2 = render_stars + draw_star_diamond (render.rs:149-235)
// =================================================================================================
// The host keeps the part that needs libm and the LCG (direction, twinkle brightness -> colour); the device does
// `dir * 10000.0` -> perspective_transform -> `cam_space.z > 0.1` -> project -> the diamond of up to nine set_pixel
// calls.  Stars are drawn in order and later ones overwrite earlier ones; a star's own pixels are distinct, so the
// order is resolved per pixel by the largest star index (atomicMax in `owner`, then the owner stores).
template <class Visit>
__device__ __forceinline__ void star_pixels(const b32_star& s, int32_t size, const CallParams& p, Visit visit) {
    TVert t = transform_vertex(s.dir[0] * 10000.0f, s.dir[1] * 10000.0f, s.dir[2] * 10000.0f, p, nullptr);   // p.cam_pos = 0 (:178)
    if (!(t.w > 0.1f)) return;                                                       // :180 (NaN: not drawn)
    const int32_t cx = f2i32(t.x), cy = f2i32(t.y);                                  // `screen.x as i32` (:197)
    const int32_t W = (int32_t)p.width, H = (int32_t)p.height;
    auto put = [&](int32_t x, int32_t y, float k) {                                  // set_pixel_safe (:237-241)
        if (x < 0 || y < 0 || x >= W || y >= H) return;
        uint32_t r = s.r, g = s.g, b = s.b;
        if (k != 1.0f) { r = f2u8((float)s.r * k); g = f2u8((float)s.g * k); b = f2u8((float)s.b * k); }   // :211-215, :224-228
        visit((uint32_t)y * p.width + (uint32_t)x, r | (g << 8) | (b << 16) | 0xFF000000u);
    };
    put(cx, cy, 1.0f);
    if (size >= 2) { put(cx - 1, cy, 0.7f); put(cx + 1, cy, 0.7f); put(cx, cy - 1, 0.7f); put(cx, cy + 1, 0.7f); }
    if (size >= 3) { put(cx - 2, cy, 0.4f); put(cx + 2, cy, 0.4f); put(cx, cy - 2, 0.4f); put(cx, cy + 2, 0.4f); }
}

__global__ void __launch_bounds__(128)
k_stars_claim(const b32_star* __restrict__ stars, uint32_t n, int32_t size, uint32_t* __restrict__ owner, CallParams p) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    star_pixels(stars[i], size, p, [&](uint32_t idx, uint32_t) { atomicMax(&owner[idx], i + 1); });
}

__global__ void __launch_bounds__(128)
k_stars_write(const b32_star* __restrict__ stars, uint32_t n, int32_t size, const uint32_t* __restrict__ owner,
              uint32_t* __restrict__ fb_rgba, CallParams p) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    star_pixels(stars[i], size, p, [&](uint32_t idx, uint32_t v) { if (owner[idx] == i + 1) fb_rgba[idx] = v; });
}

// =================================================================================================
// placed asset parts: the per-object vertex transform of render_asset_parts (src/scene.rs:141-160)
// =================================================================================================
__global__ void __launch_bounds__(256)
k_place(const b32_vertex* __restrict__ in, b32_vertex* __restrict__ out, uint32_t nv, float cos_f, float sin_f, float wx, float wy, float wz) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
        const float* v = reinterpret_cast<const float*>(in + i);
        float* o = reinterpret_cast<float*>(out + i);
        float rx = v[0] * cos_f - v[2] * sin_f;                     // rotate around Y by facing (:143-144)
        float rz = v[0] * sin_f + v[2] * cos_f;
        o[0] = rx + wx; o[1] = v[1] + wy; o[2] = rz + wz;           // then translate (:146)
        o[3] = v[3]; o[4] = v[4];                                   // uv
        o[5] = v[5] * cos_f - v[7] * sin_f;                         // normal (:148-152)
        o[6] = v[6];
        o[7] = v[5] * sin_f + v[7] * cos_f;
        o[8] = v[8];                                                // colour + blend tag
    }
}

// =================================================================================================
// small utility kernels
// =================================================================================================
__global__ void k_fb_clear(uint32_t* __restrict__ rgba, float* __restrict__ z, uint32_t n, uint32_t color) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { rgba[i] = color; z[i] = 3.40282347e+38f; }
}

// index -> CLUT expansion of a texture into the RGB555 texel pool (Clut::lookup, types.rs:390-397;
// IndexedAtlas::to_texture15, mesh_editor.rs:669-682)
__global__ void k_tex_expand(const uint8_t* __restrict__ idx, const uint16_t* __restrict__ clut, uint32_t clut_len,
                             uint32_t format, uint32_t n, uint16_t* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t k = format == B32_TEX_IDX8 ? idx[i] : ((idx[i >> 1] >> ((i & 1) * 4)) & 0xF);
        out[i] = k < clut_len ? clut[k] : (uint16_t)0;
    }
}

// "texel writes" mask of the texel pool: bit i = (texel i & 0x7FFF) != 0, i.e. the texel of a black-keyed
// surface is neither the transparent key 0x0000 nor black (render.rs:1591-1607).  One thread per word.
__global__ void k_tex_mask(const uint16_t* __restrict__ texels, uint32_t n_texels, uint32_t n_words, uint32_t* __restrict__ mask) {
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x) {
        uint32_t m = 0;
        for (uint32_t b = 0; b < 32; ++b) {
            uint32_t i = w * 32 + b;
            if (i < n_texels && (texels[i] & 0x7FFFu)) m |= 1u << b;
        }
        mask[w] = m;
    }
}

// The same for the RGB888 texel pool (one Color = r | g << 8 | b << 16 | blend << 24 per texel): a texel writes
// unless its tag is Erase (Color::is_transparent, types.rs:783-785).
__global__ void k_tex8_mask(const uint32_t* __restrict__ texels, uint32_t n_texels, uint32_t n_words, uint32_t* __restrict__ mask) {
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x) {
        uint32_t m = 0;
        for (uint32_t b = 0; b < 32; ++b) {
            uint32_t i = w * 32 + b;
            if (i < n_texels && (texels[i] >> 24) != B32_BLEND_ERASE) m |= 1u << b;
        }
        mask[w] = m;
    }
}

// TexDev.blend of an RGB888 texture = 1 iff some texel carries a PS1 blend tag (neither Opaque nor Erase): only
// surfaces of such textures (or with editor alpha) ever read the framebuffer.  blockIdx.y = texture.
__global__ void k_tex8_flags(const uint32_t* __restrict__ texels, TexDev* __restrict__ desc) {
    TexDev d = desc[blockIdx.y];
    uint32_t n = d.w * d.h;
    bool any = false;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t tag = texels[d.off + i] >> 24;
        any = any || (tag != B32_BLEND_OPAQUE && tag != B32_BLEND_ERASE);
    }
    if (__any_sync(0xFFFFFFFFu, any) && (threadIdx.x & 31) == 0) atomicOr(&desc[blockIdx.y].blend, 1u);
}

// =================================================================================================
// launchers (host)
// =================================================================================================
// Launch a frame kernel, optionally with the programmatic-stream-serialization attribute (the kernel may start
// before the previous kernel on the stream has finished; it orders itself with pdl_wait()).  With L.patch set
// the arguments go into the matching kernel node of an instantiated graph instead (see GraphPatch).
template <typename... KArgs, typename... Args, size_t... I>
static void launch_k_impl(const LaunchCtx& L, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, bool pdl,
                          std::index_sequence<I...>, Args&&... args) {
    std::tuple<KArgs...> vals(static_cast<KArgs>(args)...);
    void* ptrs[] = {static_cast<void*>(&std::get<I>(vals))...};
    ++*L.launches;
    if (L.patch) {
        GraphPatch& g = *L.patch;
        for (int i = 0; i < g.n; ++i)
            if (g.func[i] == reinterpret_cast<void*>(kern) && !g.used[i]) {
                cudaKernelNodeParams np{};
                np.func = reinterpret_cast<void*>(kern); np.gridDim = grid; np.blockDim = block;
                np.sharedMemBytes = (unsigned)smem; np.kernelParams = ptrs; np.extra = nullptr;
                cudaError_t e = cudaGraphExecKernelNodeSetParams(g.exec, g.node[i], &np);
                if (e != cudaSuccess) g.err = e;
                g.used[i] = true;
                return;
            }
        g.err = cudaErrorInvalidValue;           // the graph has no node for this kernel: topology changed
        return;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = L.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
    cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kern), ptrs);
}
template <typename... KArgs, typename... Args>
static void launch_k(const LaunchCtx& L, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, bool pdl, Args&&... args) {
    launch_k_impl(L, kern, grid, block, smem, pdl, std::index_sequence_for<KArgs...>{}, std::forward<Args>(args)...);
}

// Per-device kernel attributes (dynamic shared memory above the 48 KB default); called once per context, on its device.
int init_kernel_attributes() {
    cudaError_t e = cudaSuccess;
    auto smem_attr = [&](auto kern, size_t bytes) { if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); };
    smem_attr(k_fill_opaque<false, OpDense, false>, OpDense::SMEM);  smem_attr(k_fill_opaque<true, OpDense, false>, OpDense::SMEM);
    smem_attr(k_fill_opaque<false, OpDense, true>, OpDense::SMEM);   smem_attr(k_fill_opaque<true, OpDense, true>, OpDense::SMEM);
    smem_attr(k_fill_opaque<false, OpSparse, false>, OpSparse::SMEM); smem_attr(k_fill_opaque<true, OpSparse, false>, OpSparse::SMEM);
    smem_attr(k_fill_opaque<false, OpSparse, true>, OpSparse::SMEM);  smem_attr(k_fill_opaque<true, OpSparse, true>, OpSparse::SMEM);
    smem_attr(k_fill_ordered<false, false>, ORD_SMEM); smem_attr(k_fill_ordered<true, false>, ORD_SMEM);
    smem_attr(k_fill_ordered<false, true>, ORD_SMEM_PRE); smem_attr(k_fill_ordered<true, true>, ORD_SMEM_PRE);
    return (int)e;
}

static inline uint32_t grid_for(uint32_t n, uint32_t block, uint32_t sms, uint32_t per_sm = 8) {
    uint32_t g = (n + block - 1) / block;
    uint32_t cap = sms * per_sm;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

void launch_transform(const LaunchCtx& L, const b32_vertex* verts, TVert* out, float* dbg_cam, const CallParams& p) {
    if (p.nv == 0) return;
    launch_k(L, k_transform, grid_for(p.nv, 256, L.sms), 256, 0, false, verts, out, dbg_cam, p);
}

void launch_setup(const LaunchCtx& L, const b32_vertex* verts, const b32_face* faces, const TVert* tv, const TexDev* tex,
                  const LightDev* lights, SurfRec* recs, uint64_t* keys, BinHead* heads, uint4* masks, uint32_t* ocount,
                  WireTri* wire, CallState* st, uint32_t* zero_next, uint32_t zero_words,
                  uint32_t* clear_rgba, float* clear_z, uint32_t clear_n, uint32_t clear_color, const CallParams& p) {
    if (p.nf == 0) return;
    const size_t smem = (size_t)p.mtiles_x * p.mtiles_y * sizeof(uint4) + SETUP_STAGE_BYTES;
    launch_k(L, p.has_spot ? k_setup<true> : k_setup<false>, grid_for(std::max(p.nf, clear_n / 4), SETUP_THREADS, L.sms, 16), SETUP_THREADS, smem, tv != nullptr, verts, faces, tv, tex, lights, recs, keys, heads,
             masks, ocount, wire, st, zero_next, zero_words, clear_rgba, clear_z, clear_n, clear_color, p);
}

void launch_fill_opaque(const LaunchCtx& L, const SurfRec* recs, const uint4* masks, const BinHead* heads,
                        const TexDev* tex, const uint16_t* texels, const uint32_t* texmask, uint32_t* fb_rgba, float* fb_z,
                        CallState* st, uint32_t* sticky, BinHead* crowd, uint32_t crowd_cap, const CallParams& p) {
    uint32_t ntiles = p.tiles_x * p.tiles_y;
    if (ntiles == 0 || p.nf == 0) return;
    // The 256-thread shape packs 4 CTAs per SM: it wins when the frame has more tiles than one wave of the two-lanes-per-
    // pixel shape holds (3 CTAs per SM), and for enqueue-only calls, whose kernels share the SMs with the frames queued
    // around them.  A blocking call of up to 444 tiles has the GPU to itself: the two-lane shape finishes it sooner.
    static const bool force_dense = getenv("B32_FILL_DENSE") != nullptr, force_sparse = getenv("B32_FILL_SPARSE") != nullptr;
    const bool sparse = force_sparse || (!force_dense && (p.async_call || ntiles * OpDense::SPLIT > L.sms * (uint32_t)OpDense::MINB));
    // Float / ortho projection: no surface has integer edge values, all of them replay the reference's rounded additions:
    // the shared-edge-prefix instantiation (see k_fill_opaque).  Fixed-point calls keep the per-pixel replay for the few
    // surfaces that leave the exact-integer range (far off-screen vertices) — unless the previous call on this context
    // counted large ones (CallState.n_big_stepped -> CallParams.prefer_prefix): a camera next to a wall stays there.
    static const bool no_prefix = getenv("B32_NO_EDGE_PREFIX") != nullptr, force_prefix = getenv("B32_FORCE_EDGE_PREFIX") != nullptr;
    const bool pre = (fill_uses_edge_prefix(p) || p.prefer_prefix || force_prefix) && !no_prefix;
    using Kern = void (*)(const SurfRec*, const uint4*, const BinHead*, const TexDev*, const uint16_t*, const uint32_t*, uint32_t*, float*,
                          CallState*, uint32_t*, BinHead*, uint32_t, CallParams);
    Kern k;
    if (sparse) k = p.rgb888 ? (pre ? k_fill_opaque<true, OpSparse, true> : k_fill_opaque<true, OpSparse, false>)
                             : (pre ? k_fill_opaque<false, OpSparse, true> : k_fill_opaque<false, OpSparse, false>);
    else        k = p.rgb888 ? (pre ? k_fill_opaque<true, OpDense, true> : k_fill_opaque<true, OpDense, false>)
                             : (pre ? k_fill_opaque<false, OpDense, true> : k_fill_opaque<false, OpDense, false>);
    if (sparse) launch_k(L, k, ntiles * OpSparse::SPLIT, OpSparse::THREADS, OpSparse::SMEM, true,
                         recs, masks, heads, tex, texels, texmask, fb_rgba, fb_z, st, sticky, crowd, crowd_cap, p);
    else        launch_k(L, k, ntiles * OpDense::SPLIT, OpDense::THREADS, OpDense::SMEM, true,
                         recs, masks, heads, tex, texels, texmask, fb_rgba, fb_z, st, sticky, crowd, crowd_cap, p);
}

void launch_fill_ordered(const LaunchCtx& L, const SurfRec* recs, const uint4* masks, BinHead* scratch, const uint64_t* keys,
                         const TexDev* tex, const uint16_t* texels, uint32_t* fb_rgba, float* fb_z,
                         CallState* st, uint32_t* sticky, const CallParams& p, uint32_t scratch_cap) {
    uint32_t ntiles = p.tiles_x * p.tiles_y;
    if (ntiles == 0 || p.nf == 0) return;
    static const bool no_prefix = getenv("B32_NO_EDGE_PREFIX") != nullptr;
    const bool pre = (fill_uses_edge_prefix(p) || p.prefer_prefix) && !no_prefix;
    launch_k(L, p.rgb888 ? (pre ? k_fill_ordered<true, true> : k_fill_ordered<true, false>) : (pre ? k_fill_ordered<false, true> : k_fill_ordered<false, false>),
             ntiles, FILL_THREADS, pre ? ORD_SMEM_PRE : ORD_SMEM, p.enq_ordered != 0,
             recs, masks, scratch, keys, tex, static_cast<const void*>(texels), fb_rgba, fb_z, st, sticky, p, scratch_cap);
}

void launch_wire(const LaunchCtx& L, const WireTri* wire, uint32_t kind, uint32_t color, bool depth_test, uint32_t* table, uint32_t table_size,
                 uint32_t* fb_rgba, const float* fb_z, CallState* st, const CallParams& p) {
    if (p.nf == 0) return;
    cudaMemsetAsync(table, 0xFF, (size_t)table_size * sizeof(uint32_t), L.stream);          // WIRE_EMPTY
    k_wire_dedup<<<grid_for(p.nf * 3, 128, L.sms, 16), 128, 0, L.stream>>>(wire, p.nf, kind, table, table_size - 1);
    k_wire<<<grid_for(p.nf * 3, 128, L.sms, 16), 128, 0, L.stream>>>(wire, p.nf, kind, color, depth_test, table, table_size - 1, fb_rgba, fb_z, st, p);
    *L.launches += 2;
}

void launch_fb_clear(const LaunchCtx& L, uint32_t* rgba, float* z, uint32_t n, uint32_t color) {
    if (n == 0) return;
    launch_k(L, k_fb_clear, grid_for(n, 256, L.sms), 256, 0, false, rgba, z, n, color);
}

void launch_tex_expand(const LaunchCtx& L, const uint8_t* idx, const uint16_t* clut, uint32_t clut_len, uint32_t format, uint32_t n, uint16_t* out) {
    if (n == 0) return;
    k_tex_expand<<<grid_for(n, 256, L.sms), 256, 0, L.stream>>>(idx, clut, clut_len, format, n, out);
    ++*L.launches;
}

void launch_sky(const LaunchCtx& L, const b32_sky_vertex* verts, const uint32_t* faces, SkyRec* recs, BinHead* heads, uint4* masks,
                uint32_t* fb_rgba, CallState* st, uint32_t* zero_next, uint32_t zero_words, const CallParams& p) {
    if (p.nf == 0) return;
    const size_t smem = (size_t)p.mtiles_x * p.mtiles_y * sizeof(uint4);
    launch_k(L, k_sky_setup, grid_for(p.nf, SETUP_GROUP, L.sms, 16), SETUP_GROUP, smem, false, verts, faces, recs, heads, masks, st, zero_next, zero_words, p);
    launch_k(L, k_sky_fill, p.tiles_x * p.tiles_y, FILL_THREADS, 0, true, recs, masks, heads, fb_rgba, st, p);
}

void launch_stars(const LaunchCtx& L, const b32_star* stars, uint32_t n, int32_t size, uint32_t* owner, uint32_t* fb_rgba, const CallParams& p) {
    if (n == 0) return;
    cudaMemsetAsync(owner, 0, (size_t)p.width * p.height * 4, L.stream);
    k_stars_claim<<<(n + 127) / 128, 128, 0, L.stream>>>(stars, n, size, owner, p);
    k_stars_write<<<(n + 127) / 128, 128, 0, L.stream>>>(stars, n, size, owner, fb_rgba, p);
    *L.launches += 2;
}

void launch_place(const LaunchCtx& L, const b32_vertex* in, b32_vertex* out, uint32_t nv, float cos_f, float sin_f, const float* world_pos) {
    if (nv == 0) return;
    k_place<<<grid_for(nv, 256, L.sms), 256, 0, L.stream>>>(in, out, nv, cos_f, sin_f, world_pos[0], world_pos[1], world_pos[2]);
    ++*L.launches;
}

void launch_tex_mask(const LaunchCtx& L, const uint16_t* texels, uint32_t n_texels, uint32_t n_words, uint32_t* mask) {
    if (n_words == 0) return;
    k_tex_mask<<<grid_for(n_words, 256, L.sms), 256, 0, L.stream>>>(texels, n_texels, n_words, mask);
    ++*L.launches;
}

void launch_tex8_scan(const LaunchCtx& L, const uint32_t* texels, uint32_t n_texels, uint32_t n_words, uint32_t* mask,
                      TexDev* desc, uint32_t ntex, uint32_t max_texels) {
    if (n_words) { k_tex8_mask<<<grid_for(n_words, 256, L.sms), 256, 0, L.stream>>>(texels, n_texels, n_words, mask); ++*L.launches; }
    if (ntex && max_texels) {
        dim3 grid(std::min<uint32_t>((max_texels + 255) / 256, 64), ntex);
        k_tex8_flags<<<grid, 256, 0, L.stream>>>(texels, desc);
        ++*L.launches;
    }
}

}  // namespace b32
