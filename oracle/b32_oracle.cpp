// b32_oracle.cpp — CPU oracle (TEST INFRASTRUCTURE ONLY; see b32_oracle.h).
//
// Line-by-line C++17 restatement of the reference Rust rasterizer.  Citations are
// path:line under /root/reference/.  Build with
//     g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -shared
// (-ffp-contract=off: Rust never fuses a*b+c; every f32 operator rounds once, left to right).
//
// PARITY UNPINNED by reference-executed vectors (no rustc here, and render.rs has no tests); pinned
// by fixed.rs:477-548 facts, spec constants and agreement with oracle/pymodel.py (see header).
#include "b32_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <limits>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// Rust scalar semantics
// ---------------------------------------------------------------------------------------------
// `f as i32`: saturating, NaN -> 0 (Rust reference: "Casting", float-to-int).
inline int32_t f2i32(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
// `f as usize` on a 64-bit target.
inline uint64_t f2usize(float f) {
    if (f != f) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)f;
}
// `f as u8`
inline uint8_t f2u8(float f) {
    if (f != f) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 255.0f) return 255;
    return (uint8_t)f;
}
// f32::min / f32::max: IEEE minNum/maxNum (a NaN operand yields the other one).
inline float rmin(float a, float b) { if (a != a) return b; if (b != b) return a; return a < b ? a : b; }
inline float rmax(float a, float b) { if (a != a) return b; if (b != b) return a; return a > b ? a : b; }
// f32::clamp: NaN propagates.
inline float rclamp(float x, float lo, float hi) { if (x < lo) x = lo; if (x > hi) x = hi; return x; }
// f32::rem_euclid (core::f32): r = self % rhs; if r < 0 { r + rhs.abs() } else { r }
inline float rem_euclid(float a, float b) { float r = std::fmod(a, b); return r < 0.0f ? r + std::fabs(b) : r; }

struct V3 { float x, y, z; };
inline V3 mk3(const float* p) { return V3{p[0], p[1], p[2]}; }
inline V3 add(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }          // math.rs:60-69
inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }          // math.rs:71-80
inline V3 scale(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }           // math.rs:51-57
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }         // math.rs:23-25
inline float len(V3 a) { return std::sqrt(dot(a, a)); }                            // math.rs:35-37
inline V3 normalize(V3 a) {                                                        // math.rs:39-49
    float l = len(a);
    if (l == 0.0f) return V3{0.0f, 0.0f, 0.0f};
    return V3{a.x / l, a.y / l, a.z / l};
}
// math.rs:103-109
inline V3 perspective_transform(V3 v, V3 cx, V3 cy, V3 cz) { return V3{dot(v, cx), dot(v, cy), dot(v, cz)}; }

// math.rs:117-136
inline V3 project(V3 v, uint32_t width, uint32_t height) {
    const float ud = 5.0f;
    const float us = ud - 1.0f;
    const float vs = ((float)std::min(width, height) / 2.0f) * 0.75f;
    float denom = v.z + ud;
    if (std::fabs(denom) < 0.001f) return V3{(float)width / 2.0f, (float)height / 2.0f, v.z};
    return V3{(v.x * us) / denom * vs + ((float)width / 2.0f),
              (v.y * us) / denom * vs + ((float)height / 2.0f),
              denom};
}
// math.rs:140-148
inline V3 project_ortho(V3 v, float zoom, float cx, float cy, uint32_t width, uint32_t height) {
    return V3{(v.x - cx) * zoom + ((float)width / 2.0f),
              -(v.y - cy) * zoom + ((float)height / 2.0f),
              v.z};
}
constexpr float NEAR_PLANE = 0.1f;   // math.rs:155

// ---------------------------------------------------------------------------------------------
// fixed.rs
// ---------------------------------------------------------------------------------------------
struct UnrTable {
    uint8_t t[257];
    UnrTable() {                                   // fixed.rs:20-31
        for (uint32_t i = 0; i < 257; ++i) {
            uint32_t div = i + 256;
            uint32_t quotient = 262144u / div;
            int32_t val = (int32_t)((quotient + 1) / 2) - 257;
            t[i] = val > 0 ? (uint8_t)val : 0;
        }
    }
};
const UnrTable UNR;

inline int32_t fx_from_f32(float f) { return f2i32(f * 4096.0f); }              // fixed.rs:125-127
inline int32_t fx_from_int(int32_t n) { return (int32_t)((uint32_t)n << 12); }  // fixed.rs:119-121
inline int32_t fx_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); } // :236-238
inline int32_t fx_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); } // :244-246
inline int32_t fx_mul(int32_t a, int32_t b) {                                   // fixed.rs:161-165
    int64_t r = ((int64_t)a * (int64_t)b) >> 12;
    return (int32_t)(uint32_t)(uint64_t)r;     // `as i32` truncates
}
inline int32_t fx_floor(int32_t a) { return a >> 12; }                          // fixed.rs:137-139

int32_t fx_div_unr(int32_t self, int32_t divisor) {                             // fixed.rs:178-230
    if (divisor == 0) return 0;
    bool result_negative = (self < 0) != (divisor < 0);
    uint64_t num = (uint64_t)(self < 0 ? (uint32_t)0 - (uint32_t)self : (uint32_t)self);  // unsigned_abs
    uint32_t den = divisor < 0 ? (uint32_t)0 - (uint32_t)divisor : (uint32_t)divisor;
    if (den == 0) return 0;
    uint32_t z = (uint32_t)__builtin_clz(den);
    uint64_t d_norm = (uint64_t)den << z;
    uint64_t d16 = d_norm >> 16;
    uint64_t table_idx = std::min<uint64_t>((d16 - 0x7FC0ull) >> 7, 256);
    uint64_t u_val = (uint64_t)UNR.t[table_idx] + 0x101;
    uint64_t nr1 = (0x2000080ull - d16 * u_val) >> 8;
    uint64_t nr2 = (0x80ull + nr1 * u_val) >> 8;
    uint64_t raw = num * nr2;                       // wrapping_mul
    uint32_t shift = 36u - z;                       // wrapping_sub; z in 0..31 so 5..36
    uint64_t magnitude;
    if (shift < 64) {
        uint64_t rounding = shift > 0 ? (1ull << (shift - 1)) : 0;
        magnitude = (raw + rounding) >> shift;
    } else {
        magnitude = 0;
    }
    int32_t clamped = (int32_t)std::min<uint64_t>(magnitude, (uint64_t)INT32_MAX);
    return result_negative ? -clamped : clamped;
}

// ---------------------------------------------------------------------------------------------
// Compatibility switches — used ONLY to pin this restatement against the reference's own compiled
// code, /root/reference/docs/bonnie-32.wasm (crate version string 0.1.8; the source tree is 0.1.11).
// The binary predates four source changes; each switch turns ONE of them back so that the rest of
// the restatement can be compared bit for bit with reference-executed results (oracle/wasm/,
// tests/golden/ref_wasm/).  Default 0 = the 0.1.11 source.  Evidence for every
// item is the decompiled function (oracle/wasm/wasmdecomp.py), quoted in oracle/wasm/DRIFT.md.
enum : uint32_t {
    COMPAT_DIV_EXACT       = 1,  // project_to_screen: `((c*4) << 12) / denom`, clamped to +-2^27 and `& !7`,
                                 //   instead of 0.1.11's div_unr (fixed.rs:178-230 did not exist yet)
    COMPAT_ALWAYS_DITHER   = 2,  // rasterize_triangle_15: dither iff settings.dithering (no needs_dither rule,
                                 //   render.rs:1487-1492 is newer)
    COMPAT_RGBA_SHL3       = 4,  // Color15::r8/g8/b8 = v << 3 (0.1.11: (v<<3)|(v>>2), types.rs:138-152)
    COMPAT_TRANSP_TEX_ONLY = 8,  // has_transparency: texture present => its blend != Opaque decides alone;
                                 //   no texture => face blend decides (0.1.11: either, render.rs:2403-2415)
};
uint32_t g_compat = 0;

// 0.1.8 projection divide, from the decompiled project_fixed:
//   q = (((n << 2) as i64) << 12) / denom;  q = max(q, -134217728); q = min(q, 134213632); q &= -8
inline int32_t fx_div_exact_018(int32_t scaled_num, int32_t denom) {
    int64_t q = ((int64_t)scaled_num << 12) / (int64_t)denom;
    int32_t r = (int32_t)(uint32_t)(uint64_t)q;          // i32.wrap_i64
    r = r > -134217728 ? r : -134217728;
    r = r < 134213632 ? r : 134213632;
    return r & -8;
}

struct FxV3 { int32_t x, y, z; };
inline FxV3 fx_from_vec3(V3 v) { return FxV3{fx_from_f32(v.x), fx_from_f32(v.y), fx_from_f32(v.z)}; } // :291-297
inline int32_t fx_dot(FxV3 a, FxV3 b) {                                         // fixed.rs:311-313
    return fx_add(fx_add(fx_mul(a.x, b.x), fx_mul(a.y, b.y)), fx_mul(a.z, b.z));
}
// fixed.rs:362-381
inline FxV3 transform_to_camera_space(V3 world, V3 cam_pos, V3 bx, V3 by, V3 bz) {
    FxV3 w = fx_from_vec3(world), c = fx_from_vec3(cam_pos);
    FxV3 rel{fx_sub(w.x, c.x), fx_sub(w.y, c.y), fx_sub(w.z, c.z)};
    FxV3 fbx = fx_from_vec3(bx), fby = fx_from_vec3(by), fbz = fx_from_vec3(bz);
    return FxV3{fx_dot(rel, fbx), fx_dot(rel, fby), fx_dot(rel, fbz)};
}
// fixed.rs:390-420
inline void project_to_screen(FxV3 cam, uint32_t width, uint32_t height, int32_t* sx, int32_t* sy, int32_t* depth) {
    int32_t distance = fx_from_f32(5.0f);
    int32_t scl = fx_from_f32(4.0f);
    int32_t viewport_scale = fx_from_f32(((float)std::min(width, height) / 2.0f) * 0.75f);
    int32_t half_w = fx_from_int((int32_t)width / 2);
    int32_t half_h = fx_from_int((int32_t)height / 2);
    int32_t denom = fx_add(cam.z, distance);
    // i32::abs in a release build wraps for i32::MIN (stays negative => "< 256" holds)
    int32_t adenom = denom < 0 ? (int32_t)((uint32_t)0 - (uint32_t)denom) : denom;
    if (adenom < 256) { *sx = fx_floor(half_w); *sy = fx_floor(half_h); *depth = cam.z; return; }
    int32_t proj_x, proj_y;
    if (g_compat & COMPAT_DIV_EXACT) {
        // the binary computes `cam << 2` (wrapping) where 0.1.11 has cam * from_f32(4.0)
        proj_x = fx_div_exact_018((int32_t)((uint32_t)cam.x << 2), denom);
        proj_y = fx_div_exact_018((int32_t)((uint32_t)cam.y << 2), denom);
    } else {
        proj_x = fx_div_unr(fx_mul(cam.x, scl), denom);
        proj_y = fx_div_unr(fx_mul(cam.y, scl), denom);
    }
    int32_t screen_x = fx_add(fx_mul(proj_x, viewport_scale), half_w);
    int32_t screen_y = fx_add(fx_mul(proj_y, viewport_scale), half_h);
    *sx = fx_floor(screen_x); *sy = fx_floor(screen_y); *depth = cam.z;
}
// fixed.rs:424-441
inline void project_fixed(V3 world, const b32_camera* cam, uint32_t w, uint32_t h, int32_t* sx, int32_t* sy, float* depth) {
    FxV3 c = transform_to_camera_space(world, mk3(cam->position), mk3(cam->basis_x), mk3(cam->basis_y), mk3(cam->basis_z));
    int32_t d;
    project_to_screen(c, w, h, sx, sy, &d);
    *depth = (float)d / 4096.0f;                                                // fixed.rs:131-133
}

// ---------------------------------------------------------------------------------------------
// types.rs: Color15, Texture15, Clut
// ---------------------------------------------------------------------------------------------
inline uint8_t c15_r5(uint16_t c) { return (uint8_t)((c >> 10) & 0x1F); }      // types.rs:121-136
inline uint8_t c15_g5(uint16_t c) { return (uint8_t)((c >> 5) & 0x1F); }
inline uint8_t c15_b5(uint16_t c) { return (uint8_t)(c & 0x1F); }
inline uint8_t expand_5_to_8(uint8_t v) { return (uint8_t)((v << 3) | (v >> 2)); } // render.rs:1161-1163
inline uint16_t c15_new_semi(uint8_t r, uint8_t g, uint8_t b, bool semi) {      // types.rs:41-56
    uint16_t c = (uint16_t)(((uint16_t)std::min<uint8_t>(r, 31) << 10) | ((uint16_t)std::min<uint8_t>(g, 31) << 5) |
                            (uint16_t)std::min<uint8_t>(b, 31));
    if (semi) c |= 0x8000;
    return c;
}

// A Texture15 as the reference rasterizer sees it (types.rs:532-539).
struct Tex15 {
    uint32_t width = 0, height = 0;
    std::vector<uint16_t> pixels;
    uint32_t blend_mode = 0;
};
// Clut::lookup (types.rs:390-397) applied per texel = IndexedAtlas::to_texture15
// (src/modeler/mesh_editor.rs:669-682); RGB555 textures are taken as they are.
Tex15 to_texture15(const b32_tex_desc& d) {
    Tex15 t;
    t.width = d.width; t.height = d.height; t.blend_mode = d.blend_mode;
    size_t n = (size_t)d.width * d.height;
    t.pixels.resize(n);
    if (d.format == B32_TEX_RGB555) {
        if (n) std::memcpy(t.pixels.data(), d.pixels, n * 2);
    } else {
        const uint8_t* idx = (const uint8_t*)d.pixels;
        for (size_t i = 0; i < n; ++i) {
            uint8_t k = d.format == B32_TEX_IDX8 ? idx[i] : (uint8_t)((idx[i >> 1] >> ((i & 1) * 4)) & 0xF);
            t.pixels[i] = (k < d.clut_len) ? d.clut[k] : (uint16_t)0x0000;
        }
    }
    return t;
}
// Texture15::sample, types.rs:671-681
inline uint16_t tex_sample(const Tex15& t, float u, float v) {
    if (t.width == 0 || t.height == 0 || t.pixels.empty()) return 0x0000;
    float u_wrapped = rem_euclid(u, 1.0f);
    float v_wrapped = rem_euclid(v, 1.0f);
    uint64_t tx = std::min<uint64_t>(f2usize(u_wrapped * (float)t.width), t.width - 1);
    uint64_t ty = std::min<uint64_t>(f2usize(v_wrapped * (float)t.height), t.height - 1);
    return t.pixels[ty * t.width + tx];
}

// ---------------------------------------------------------------------------------------------
// render.rs helpers
// ---------------------------------------------------------------------------------------------
struct Col { uint8_t r, g, b, blend; };   // types.rs:719-726
inline bool col_eq(Col a, Col b) { return a.r == b.r && a.g == b.g && a.b == b.b && a.blend == b.blend; }

// f32::acos of the reference's shipped build.  Rust lowers `f32::acos` to the `acosf` symbol: on wasm32 that is
// compiler_builtins' `libm` crate (a port of musl / FreeBSD e_acosf.c), on native targets the platform libm (which may
// differ in the last ulp).  This restates the function as it stands in the reference's own binary
// (docs/bonnie-32.wasm, `compiler_builtins::math::partial_availability::acosf`, func 2057; disassembly checked op by
// op) and is pinned against it on 400 000 inputs (tests/test_ref_wasm.py::test_acosf_matches_reference_binary).
// Every operator rounds once (-ffp-contract=off); the binary drops musl's `+ 0x1p-120f` terms (they do not change the
// rounded result).
inline float acosf_rpoly(float z) {
    float p = z * (0.16666586697101593f + z * (-0.04274342209100723f + z * -0.008656363002955914f));
    float q = z * -0.7066296339035034f + 1.0f;
    return p / q;
}
float ref_acosf(float x) {
    const float pio2_hi = 1.570796251296997f;        // 0x3fc90fda
    const float pio2_lo = 7.549789415861596e-08f;    // 0x33a22168
    uint32_t hx; std::memcpy(&hx, &x, 4);
    uint32_t ix = hx & 0x7fffffffu;
    if (ix >= 0x3f800000u) {                         // |x| >= 1 or NaN
        if (ix == 0x3f800000u) return (hx >> 31) ? 3.141592502593994f : 0.0f;
        return 0.0f / (x - x);
    }
    if (ix < 0x3f000000u) {                          // |x| < 0.5
        if (ix <= 0x32800000u) return pio2_hi;       // |x| < 2^-26
        return pio2_hi - (x - (pio2_lo - x * acosf_rpoly(x * x)));
    }
    if (hx >> 31) {                                  // x < -0.5
        float z = (1.0f + x) * 0.5f;
        float s = std::sqrt(z);
        float w = acosf_rpoly(z) * s - pio2_lo;
        float t = pio2_hi - (s + w);
        return t + t;
    }
    float z = (1.0f - x) * 0.5f;                     // x > 0.5
    float s = std::sqrt(z);
    uint32_t sb; std::memcpy(&sb, &s, 4); sb &= 0xfffff000u;
    float df; std::memcpy(&df, &sb, 4);
    float c = (z - df * df) / (s + df);
    float w = acosf_rpoly(z) * s + c;
    float t = df + w;
    return t + t;
}

// render.rs:1013-1071
void shade_multi_light_color(V3 normal, V3 world_pos, const b32_light* lights, uint32_t n, float ambient, float out[3]) {
    float total_r = ambient, total_g = ambient, total_b = ambient;
    for (uint32_t i = 0; i < n; ++i) {
        const b32_light& L = lights[i];
        if (!L.enabled) continue;
        float contribution;
        if (L.type == B32_LIGHT_DIRECTIONAL) {
            V3 neg_dir = scale(mk3(L.direction), -1.0f);
            float n_dot_l = rmax(dot(normal, neg_dir), 0.0f);
            contribution = n_dot_l * L.intensity;
        } else if (L.type == B32_LIGHT_POINT) {
            V3 to_light = sub(mk3(L.position), world_pos);
            float dist = len(to_light);
            if (dist > L.radius || dist < 0.001f) {
                contribution = 0.0f;
            } else {
                float attenuation = 1.0f - (dist / L.radius);
                float n_dot_l = rmax(dot(normal, normalize(to_light)), 0.0f);
                contribution = n_dot_l * L.intensity * attenuation * attenuation;
            }
        } else {  // Spot (render.rs:1040-1058); acos = ref_acosf above
            V3 to_light = sub(mk3(L.position), world_pos);
            float dist = len(to_light);
            if (dist > L.radius || dist < 0.001f) {
                contribution = 0.0f;
            } else {
                V3 to_surface = normalize(to_light);
                V3 neg = scale(to_surface, -1.0f);
                float spot_angle = ref_acosf(dot(neg, mk3(L.direction)));
                if (spot_angle > L.angle) {
                    contribution = 0.0f;
                } else {
                    float attenuation = 1.0f - (dist / L.radius);
                    float edge_falloff = 1.0f - (spot_angle / L.angle);
                    float n_dot_l = rmax(dot(normal, to_surface), 0.0f);
                    contribution = n_dot_l * L.intensity * attenuation * attenuation * edge_falloff;
                }
            }
        }
        float light_r = (float)L.r / 255.0f, light_g = (float)L.g / 255.0f, light_b = (float)L.b / 255.0f;
        total_r += contribution * light_r;
        total_g += contribution * light_g;
        total_b += contribution * light_b;
    }
    out[0] = rmin(total_r, 1.0f); out[1] = rmin(total_g, 1.0f); out[2] = rmin(total_b, 1.0f);
}

// render.rs:1093-1145
inline void blend_rgb555(uint8_t fr, uint8_t fg, uint8_t fb, uint8_t br, uint8_t bg, uint8_t bb, uint32_t mode, uint8_t out[3]) {
    uint8_t f5[3] = {(uint8_t)(fr >> 3), (uint8_t)(fg >> 3), (uint8_t)(fb >> 3)};
    uint8_t b5[3] = {(uint8_t)(br >> 3), (uint8_t)(bg >> 3), (uint8_t)(bb >> 3)};
    for (int i = 0; i < 3; ++i) {
        uint8_t r5;
        switch (mode) {
            case B32_BLEND_OPAQUE:      r5 = f5[i]; break;
            case B32_BLEND_AVERAGE:     r5 = (uint8_t)std::min<uint16_t>((uint16_t)((b5[i] + f5[i]) / 2), 31); break;
            case B32_BLEND_ADD:         r5 = (uint8_t)std::min<uint16_t>((uint16_t)(b5[i] + f5[i]), 31); break;
            case B32_BLEND_SUBTRACT:    r5 = (uint8_t)std::max<int16_t>((int16_t)((int16_t)b5[i] - (int16_t)f5[i]), 0); break;
            case B32_BLEND_ADD_QUARTER: r5 = (uint8_t)std::min<uint16_t>((uint16_t)(b5[i] + f5[i] / 4), 31); break;
            default:                    r5 = b5[i]; break;   // Erase
        }
        out[i] = (uint8_t)(r5 << 3);
    }
}

// render.rs:1150-1155
const int8_t PS1_DITHER_MATRIX[4][4] = {{-4, 0, -3, 1}, {2, -2, 3, -1}, {-3, 1, -4, 0}, {3, -1, 2, -2}};
// render.rs:1173-1182
inline void dither_and_quantize(uint8_t r8, uint8_t g8, uint8_t b8, uint64_t x, uint64_t y, uint8_t out[3]) {
    int32_t offset = PS1_DITHER_MATRIX[y & 3][x & 3];
    out[0] = (uint8_t)std::clamp(((int32_t)r8 + offset) >> 3, 0, 31);
    out[1] = (uint8_t)std::clamp(((int32_t)g8 + offset) >> 3, 0, 31);
    out[2] = (uint8_t)std::clamp(((int32_t)b8 + offset) >> 3, 0, 31);
}

struct Fb {                                  // render.rs:10-15
    uint8_t* pixels; float* zbuffer; uint64_t width, height;
};
// Color::to_bytes, types.rs:829-832
inline void fb_set_pixel(Fb& fb, uint64_t x, uint64_t y, Col c) {       // render.rs:301-310
    if (x < fb.width && y < fb.height) {
        uint64_t idx = (y * fb.width + x) * 4;
        fb.pixels[idx] = c.r; fb.pixels[idx + 1] = c.g; fb.pixels[idx + 2] = c.b;
        fb.pixels[idx + 3] = c.blend == B32_BLEND_ERASE ? 0 : 255;
    }
}
// Color15::r8/g8/b8, types.rs:138-152
inline uint8_t c15_ch8(uint8_t v5) { return (g_compat & COMPAT_RGBA_SHL3) ? (uint8_t)(v5 << 3) : expand_5_to_8(v5); }
// Color15::to_rgba, types.rs:220-226
inline void c15_to_rgba(uint16_t c, uint8_t out[4]) {
    if (c == 0x0000) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    out[0] = c15_ch8(c15_r5(c)); out[1] = c15_ch8(c15_g5(c)); out[2] = c15_ch8(c15_b5(c)); out[3] = 255;
}
inline void fb_set_pixel_15(Fb& fb, uint64_t x, uint64_t y, uint16_t c) {           // render.rs:445-454
    if (x < fb.width && y < fb.height) {
        uint64_t idx = (y * fb.width + x) * 4;
        c15_to_rgba(c, &fb.pixels[idx]);
    }
}
inline void fb_set_pixel_blended_15(Fb& fb, uint64_t x, uint64_t y, uint16_t c, uint32_t mode) {  // render.rs:479-502
    if (x < fb.width && y < fb.height) {
        uint64_t idx = (y * fb.width + x) * 4;
        uint8_t br = fb.pixels[idx], bg = fb.pixels[idx + 1], bb = fb.pixels[idx + 2];
        uint8_t o[3];
        if (c & 0x8000) blend_rgb555(expand_5_to_8(c15_r5(c)), expand_5_to_8(c15_g5(c)), expand_5_to_8(c15_b5(c)), br, bg, bb, mode, o);
        else { o[0] = c15_ch8(c15_r5(c)); o[1] = c15_ch8(c15_g5(c)); o[2] = c15_ch8(c15_b5(c)); }
        fb.pixels[idx] = o[0]; fb.pixels[idx + 1] = o[1]; fb.pixels[idx + 2] = o[2]; fb.pixels[idx + 3] = 255;
    }
}
inline void fb_set_pixel_xray_15(Fb& fb, uint64_t x, uint64_t y, uint16_t c) {       // render.rs:507-526
    if (x < fb.width && y < fb.height) {
        uint64_t idx = (y * fb.width + x) * 4;
        uint8_t br = fb.pixels[idx], bg = fb.pixels[idx + 1], bb = fb.pixels[idx + 2];
        fb.pixels[idx]     = (uint8_t)(((uint16_t)c15_ch8(c15_r5(c)) + br) / 2);
        fb.pixels[idx + 1] = (uint8_t)(((uint16_t)c15_ch8(c15_g5(c)) + bg) / 2);
        fb.pixels[idx + 2] = (uint8_t)(((uint16_t)c15_ch8(c15_b5(c)) + bb) / 2);
        fb.pixels[idx + 3] = 255;
    }
}
// shared tail of the two editor-alpha writers, render.rs:575-594 / 609-627
inline void editor_alpha_write(Fb& fb, uint64_t idx, uint16_t c, uint32_t mode, uint8_t editor_alpha) {
    uint8_t br = fb.pixels[idx], bg = fb.pixels[idx + 1], bb = fb.pixels[idx + 2];
    uint8_t p[3];
    if ((c & 0x8000) && mode != B32_BLEND_OPAQUE)
        blend_rgb555(expand_5_to_8(c15_r5(c)), expand_5_to_8(c15_g5(c)), expand_5_to_8(c15_b5(c)), br, bg, bb, mode, p);
    else { p[0] = expand_5_to_8(c15_r5(c)); p[1] = expand_5_to_8(c15_g5(c)); p[2] = expand_5_to_8(c15_b5(c)); }
    uint16_t a = editor_alpha, inv_a = (uint16_t)(255 - a);
    fb.pixels[idx]     = (uint8_t)(((uint16_t)p[0] * a + (uint16_t)br * inv_a) / 255);
    fb.pixels[idx + 1] = (uint8_t)(((uint16_t)p[1] * a + (uint16_t)bg * inv_a) / 255);
    fb.pixels[idx + 2] = (uint8_t)(((uint16_t)p[2] * a + (uint16_t)bb * inv_a) / 255);
    fb.pixels[idx + 3] = 255;
}
inline void fb_set_pixel_with_editor_alpha_15(Fb& fb, uint64_t x, uint64_t y, uint16_t c, uint32_t mode, uint8_t ea) {  // :567-594
    if (ea == 0 || x >= fb.width || y >= fb.height) return;
    editor_alpha_write(fb, (y * fb.width + x) * 4, c, mode, ea);
}
inline bool fb_set_pixel_with_depth_and_editor_alpha_15(Fb& fb, uint64_t x, uint64_t y, float z, uint16_t c,
                                                        uint32_t mode, uint8_t ea, bool skip_z_write) {  // :598-628
    if (ea == 0 || x >= fb.width || y >= fb.height) return false;
    uint64_t depth_idx = y * fb.width + x;
    if (z >= fb.zbuffer[depth_idx]) return false;
    if (!skip_z_write) fb.zbuffer[depth_idx] = z;
    editor_alpha_write(fb, depth_idx * 4, c, mode, ea);
    return true;
}

// Bresenham lines used by the wireframe phase: render.rs:714-751 (draw_line), :768-817 (draw_line_3d)
void fb_draw_line(Fb& fb, int32_t x0, int32_t y0, int32_t x1, int32_t y1, Col color) {
    int32_t dx = std::abs(x1 - x0), dy = -std::abs(y1 - y0);
    int32_t sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
    int32_t err = dx + dy, x = x0, y = y0;
    for (;;) {
        if (x >= 0 && x < (int32_t)fb.width && y >= 0 && y < (int32_t)fb.height) fb_set_pixel(fb, (uint64_t)x, (uint64_t)y, color);
        if (x == x1 && y == y1) break;
        int32_t e2 = 2 * err;
        if (e2 >= dy) { err += dy; x += sx; }
        if (e2 <= dx) { err += dx; y += sy; }
    }
}
void fb_draw_line_3d(Fb& fb, int32_t x0, int32_t y0, float z0, int32_t x1, int32_t y1, float z1, Col color) {
    int32_t dx = std::abs(x1 - x0), dy = -std::abs(y1 - y0);
    int32_t sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
    int32_t err = dx + dy, x = x0, y = y0;
    float total_steps = (float)std::max(dx, std::max(-dy, 1));
    float step = 0.0f;
    for (;;) {
        if (x >= 0 && x < (int32_t)fb.width && y >= 0 && y < (int32_t)fb.height) {
            float t = step / total_steps;
            float z = z0 + t * (z1 - z0);
            uint64_t idx = (uint64_t)y * fb.width + (uint64_t)x;
            if (z < fb.zbuffer[idx]) fb_set_pixel(fb, (uint64_t)x, (uint64_t)y, color);
        }
        if (x == x1 && y == y1) break;
        int32_t e2 = 2 * err;
        if (e2 >= dy) { err += dy; x += sx; step += 1.0f; }
        if (e2 <= dx) { err += dx; y += sy; if (e2 < dy) step += 1.0f; }
    }
}

// struct Surface, render.rs:975-1000 (dead fields vn1-3/normal are still computed and stored so the
// CPU baseline does the reference's work).
struct Surface {
    V3 v1, v2, v3;
    V3 w1, w2, w3;
    V3 vn1, vn2, vn3;
    V3 wn1, wn2, wn3;
    float uv1[2], uv2[2], uv3[2];
    Col vc1, vc2, vc3;
    V3 normal;
    uint64_t face_idx;
    bool black_transparent;
    bool has_transparency;
    uint32_t blend_mode;
    uint8_t editor_alpha;
};


// render.rs:1440-1714
void rasterize_triangle_15(Fb& fb, const Surface& surface, const Tex15* texture, uint32_t face_blend_mode,
                           bool black_transparent, const b32_settings& settings, bool skip_z_write) {
    uint32_t blend_mode = texture ? texture->blend_mode : face_blend_mode;                    // :1450-1452

    uint64_t min_x = f2usize(rmax(rmin(rmin(surface.v1.x, surface.v2.x), surface.v3.x), 0.0f));            // :1455
    uint64_t max_x = f2usize(rmin(rmax(rmax(surface.v1.x, surface.v2.x), surface.v3.x) + 1.0f, (float)fb.width));
    uint64_t min_y = f2usize(rmax(rmin(rmin(surface.v1.y, surface.v2.y), surface.v3.y), 0.0f));
    uint64_t max_y = f2usize(rmin(rmax(rmax(surface.v1.y, surface.v2.y), surface.v3.y) + 1.0f, (float)fb.height));
    if (min_x >= max_x || min_y >= max_y) return;                                              // :1461-1463

    float flat_shade[3] = {1.0f, 1.0f, 1.0f};
    if (settings.shading == B32_SHADE_FLAT) {                                                  // :1466-1472
        V3 center_pos = scale(add(add(surface.w1, surface.w2), surface.w3), 1.0f / 3.0f);
        V3 world_normal = normalize(scale(add(add(surface.wn1, surface.wn2), surface.wn3), 1.0f / 3.0f));
        shade_multi_light_color(world_normal, center_pos, settings.lights, settings.n_lights, settings.ambient, flat_shade);
    }
    float gs1[3], gs2[3], gs3[3];
    bool gouraud = settings.shading == B32_SHADE_GOURAUD;
    if (gouraud) {                                                                             // :1475-1483
        shade_multi_light_color(surface.wn1, surface.w1, settings.lights, settings.n_lights, settings.ambient, gs1);
        shade_multi_light_color(surface.wn2, surface.w2, settings.lights, settings.n_lights, settings.ambient, gs2);
        shade_multi_light_color(surface.wn3, surface.w3, settings.lights, settings.n_lights, settings.ambient, gs3);
    }
    bool needs_dither = settings.dithering && (gouraud || texture != nullptr ||               // :1487-1492
                                               !col_eq(surface.vc1, surface.vc2) || !col_eq(surface.vc2, surface.vc3));
    if (g_compat & COMPAT_ALWAYS_DITHER) needs_dither = settings.dithering;

    V3 v1 = surface.v1, v2 = surface.v2, v3 = surface.v3;
    float area = (v2.y - v3.y) * (v1.x - v3.x) + (v3.x - v2.x) * (v1.y - v3.y);              // :1500
    if (std::fabs(area) < 0.00001f) return;
    float inv_area = 1.0f / area;
    float a0 = v2.y - v3.y, b0 = v3.x - v2.x, a1 = v3.y - v1.y, b1 = v1.x - v3.x;            // :1507-1510
    float start_x = (float)min_x, start_y = (float)min_y;
    float w0_row = a0 * (start_x - v3.x) + b0 * (start_y - v3.y);                             // :1517-1518
    float w1_row = a1 * (start_x - v3.x) + b1 * (start_y - v3.y);

    for (uint64_t y = min_y; y < max_y; ++y) {                                                 // :1530
        float w0 = w0_row, w1 = w1_row;
        for (uint64_t x = min_x; x < max_x; ++x) {
            float bc_x = w0 * inv_area;
            float bc_y = w1 * inv_area;
            float bc_z = 1.0f - bc_x - bc_y;
            const float ERR = -0.0001f;
            if (bc_x >= ERR && bc_y >= ERR && bc_z >= ERR) {                                   // :1542
                float inv_z1 = 1.0f / v1.z, inv_z2 = 1.0f / v2.z, inv_z3 = 1.0f / v3.z;
                float inv_z_interp = bc_x * inv_z1 + bc_y * inv_z2 + bc_z * inv_z3;
                float z = 1.0f / inv_z_interp;

                if (settings.use_zbuffer && !settings.xray_mode) {                             // :1553-1560
                    uint64_t idx = y * fb.width + x;
                    if (z >= fb.zbuffer[idx]) { w0 += a0; w1 += a1; continue; }
                }

                float u, v;
                if (settings.affine_textures) {                                                // :1563-1579
                    u = bc_x * surface.uv1[0] + bc_y * surface.uv2[0] + bc_z * surface.uv3[0];
                    v = bc_x * surface.uv1[1] + bc_y * surface.uv2[1] + bc_z * surface.uv3[1];
                } else {
                    float u_over_z = bc_x * surface.uv1[0] * inv_z1 + bc_y * surface.uv2[0] * inv_z2 + bc_z * surface.uv3[0] * inv_z3;
                    float v_over_z = bc_x * surface.uv1[1] * inv_z1 + bc_y * surface.uv2[1] * inv_z2 + bc_z * surface.uv3[1] * inv_z3;
                    u = u_over_z / inv_z_interp;
                    v = v_over_z / inv_z_interp;
                }

                uint16_t color = texture ? tex_sample(*texture, u, 1.0f - v) : (uint16_t)0x7FFF;  // :1582-1586

                bool is_black = c15_r5(color) == 0 && c15_g5(color) == 0 && c15_b5(color) == 0;  // :1591-1607
                if (color == 0x0000) {
                    if (is_black && !black_transparent) color = 0x8000;
                    else { w0 += a0; w1 += a1; continue; }
                } else if (black_transparent && is_black) {
                    w0 += a0; w1 += a1; continue;
                }

                uint8_t tex_r8 = expand_5_to_8(c15_r5(color));                                 // :1613-1615
                uint8_t tex_g8 = expand_5_to_8(c15_g5(color));
                uint8_t tex_b8 = expand_5_to_8(c15_b5(color));

                uint8_t vertex_r = f2u8(bc_x * (float)surface.vc1.r + bc_y * (float)surface.vc2.r + bc_z * (float)surface.vc3.r);  // :1618-1620
                uint8_t vertex_g = f2u8(bc_x * (float)surface.vc1.g + bc_y * (float)surface.vc2.g + bc_z * (float)surface.vc3.g);
                uint8_t vertex_b = f2u8(bc_x * (float)surface.vc1.b + bc_y * (float)surface.vc2.b + bc_z * (float)surface.vc3.b);

                uint8_t mod_r8 = (uint8_t)std::min<uint32_t>(((uint32_t)tex_r8 * vertex_r) / 128, 255);  // :1624-1626
                uint8_t mod_g8 = (uint8_t)std::min<uint32_t>(((uint32_t)tex_g8 * vertex_g) / 128, 255);
                uint8_t mod_b8 = (uint8_t)std::min<uint32_t>(((uint32_t)tex_b8 * vertex_b) / 128, 255);

                float shade_r, shade_g, shade_b;                                               // :1629-1640
                if (settings.shading == B32_SHADE_NONE) { shade_r = shade_g = shade_b = 1.0f; }
                else if (settings.shading == B32_SHADE_FLAT) { shade_r = flat_shade[0]; shade_g = flat_shade[1]; shade_b = flat_shade[2]; }
                else {
                    shade_r = bc_x * gs1[0] + bc_y * gs2[0] + bc_z * gs3[0];
                    shade_g = bc_x * gs1[1] + bc_y * gs2[1] + bc_z * gs3[1];
                    shade_b = bc_x * gs1[2] + bc_y * gs2[2] + bc_z * gs3[2];
                }

                uint8_t shaded_r8 = f2u8(rmin((float)mod_r8 * rclamp(shade_r, 0.0f, 2.0f), 255.0f));   // :1643-1645
                uint8_t shaded_g8 = f2u8(rmin((float)mod_g8 * rclamp(shade_g, 0.0f, 2.0f), 255.0f));
                uint8_t shaded_b8 = f2u8(rmin((float)mod_b8 * rclamp(shade_b, 0.0f, 2.0f), 255.0f));

                uint8_t q[3];                                                                  // :1649-1654
                if (needs_dither) dither_and_quantize(shaded_r8, shaded_g8, shaded_b8, x, y, q);
                else { q[0] = shaded_r8 >> 3; q[1] = shaded_g8 >> 3; q[2] = shaded_b8 >> 3; }

                bool is_all_black = q[0] == 0 && q[1] == 0 && q[2] == 0;                       // :1659-1661
                bool semi = (color & 0x8000) != 0 || is_all_black;
                uint16_t out = c15_new_semi(q[0], q[1], q[2], semi);

                uint8_t editor_alpha = surface.editor_alpha;                                   // :1664-1669
                if (editor_alpha == 0) { w0 += a0; w1 += a1; continue; }

                if (settings.xray_mode) {                                                      // :1671-1673
                    fb_set_pixel_xray_15(fb, x, y, out);
                } else if (editor_alpha < 255) {                                               // :1674-1680
                    if (settings.use_zbuffer) fb_set_pixel_with_depth_and_editor_alpha_15(fb, x, y, z, out, blend_mode, editor_alpha, skip_z_write);
                    else fb_set_pixel_with_editor_alpha_15(fb, x, y, out, blend_mode, editor_alpha);
                } else if (settings.use_zbuffer) {                                             // :1681-1694
                    uint64_t idx = y * fb.width + x;
                    if (z < fb.zbuffer[idx]) {
                        if (!skip_z_write) fb.zbuffer[idx] = z;
                        if ((out & 0x8000) && blend_mode != B32_BLEND_OPAQUE) fb_set_pixel_blended_15(fb, x, y, out, blend_mode);
                        else fb_set_pixel_15(fb, x, y, out);
                    }
                } else {                                                                       // :1695-1702
                    if ((out & 0x8000) && blend_mode != B32_BLEND_OPAQUE) fb_set_pixel_blended_15(fb, x, y, out, blend_mode);
                    else fb_set_pixel_15(fb, x, y, out);
                }
            }
            w0 += a0; w1 += a1;                                                                // :1706-1707
        }
        w0_row += b0; w1_row += b1;                                                            // :1711-1712
    }
}

// =============================================================================================
// RGB888 path: struct Texture (types.rs:1058-1066), Color ops (types.rs:783-934), writers
// (render.rs:301-437), rasterize_triangle (render.rs:1202-1433)
// =============================================================================================
struct Tex8 {
    uint32_t width = 0, height = 0;
    const Col* pixels = nullptr;       // width * height Colors (r, g, b, blend)
    uint32_t blend_mode = 0;
};
// Texture::sample, types.rs:1242-1253
inline Col tex8_sample(const Tex8& t, float u, float v) {
    if (t.width == 0 || t.height == 0 || t.pixels == nullptr) return Col{0, 0, 0, B32_BLEND_ERASE};   // Color::TRANSPARENT
    float u_wrapped = rem_euclid(u, 1.0f);
    float v_wrapped = rem_euclid(v, 1.0f);
    uint64_t tx = std::min<uint64_t>(f2usize(u_wrapped * (float)t.width), t.width - 1);
    uint64_t ty = std::min<uint64_t>(f2usize(v_wrapped * (float)t.height), t.height - 1);
    return t.pixels[ty * t.width + tx];
}
// Color::modulate, types.rs:801-808
inline Col col_modulate(Col c, Col vc) {
    return Col{(uint8_t)std::min<uint16_t>((uint16_t)((uint16_t)c.r * (uint16_t)vc.r / 128), 255),
               (uint8_t)std::min<uint16_t>((uint16_t)((uint16_t)c.g * (uint16_t)vc.g / 128), 255),
               (uint8_t)std::min<uint16_t>((uint16_t)((uint16_t)c.b * (uint16_t)vc.b / 128), 255), c.blend};
}
// shade_color_rgb, render.rs:1074-1081 (no clamp of the shade factor; `as u8` saturates, NaN -> 0)
inline Col shade_color_rgb(Col c, float sr, float sg, float sb) {
    return Col{f2u8(rmin((float)c.r * sr, 255.0f)), f2u8(rmin((float)c.g * sg, 255.0f)), f2u8(rmin((float)c.b * sb, 255.0f)), c.blend};
}
// apply_dither, render.rs:1186-1197
inline Col apply_dither(Col c, uint64_t x, uint64_t y) {
    int32_t offset = PS1_DITHER_MATRIX[y & 3][x & 3];
    uint8_t r5 = (uint8_t)std::clamp(((int32_t)c.r + offset) >> 3, 0, 31);
    uint8_t g5 = (uint8_t)std::clamp(((int32_t)c.g + offset) >> 3, 0, 31);
    uint8_t b5 = (uint8_t)std::clamp(((int32_t)c.b + offset) >> 3, 0, 31);
    return Col{(uint8_t)(r5 << 3), (uint8_t)(g5 << 3), (uint8_t)(b5 << 3), c.blend};
}
// Color::blend(back, mode) = with_blend(.., mode).blend_with(back), types.rs:886-936
inline Col col_blend(Col f, Col back, uint32_t mode) {
    switch (mode) {
        case B32_BLEND_OPAQUE: return Col{f.r, f.g, f.b, B32_BLEND_OPAQUE};
        case B32_BLEND_AVERAGE:
            return Col{(uint8_t)(((uint16_t)back.r + f.r) / 2), (uint8_t)(((uint16_t)back.g + f.g) / 2), (uint8_t)(((uint16_t)back.b + f.b) / 2), B32_BLEND_OPAQUE};
        case B32_BLEND_ADD:
            return Col{(uint8_t)std::min<uint16_t>((uint16_t)back.r + f.r, 255), (uint8_t)std::min<uint16_t>((uint16_t)back.g + f.g, 255),
                       (uint8_t)std::min<uint16_t>((uint16_t)back.b + f.b, 255), B32_BLEND_OPAQUE};
        case B32_BLEND_SUBTRACT:
            return Col{(uint8_t)std::max<int16_t>((int16_t)((int16_t)back.r - (int16_t)f.r), 0), (uint8_t)std::max<int16_t>((int16_t)((int16_t)back.g - (int16_t)f.g), 0),
                       (uint8_t)std::max<int16_t>((int16_t)((int16_t)back.b - (int16_t)f.b), 0), B32_BLEND_OPAQUE};
        case B32_BLEND_ADD_QUARTER:
            return Col{(uint8_t)std::min<uint16_t>((uint16_t)back.r + (uint16_t)f.r / 4, 255), (uint8_t)std::min<uint16_t>((uint16_t)back.g + (uint16_t)f.g / 4, 255),
                       (uint8_t)std::min<uint16_t>((uint16_t)back.b + (uint16_t)f.b / 4, 255), B32_BLEND_OPAQUE};
        default: return Col{0, 0, 0, B32_BLEND_ERASE};                                 // Color::TRANSPARENT
    }
}
inline Col fb_back(const Fb& fb, uint64_t idx) { return Col{fb.pixels[idx], fb.pixels[idx + 1], fb.pixels[idx + 2], B32_BLEND_OPAQUE}; }
inline void fb_put(Fb& fb, uint64_t idx, Col c) {                                        // Color::to_bytes, types.rs:829-832
    fb.pixels[idx] = c.r; fb.pixels[idx + 1] = c.g; fb.pixels[idx + 2] = c.b; fb.pixels[idx + 3] = c.blend == B32_BLEND_ERASE ? 0 : 255;
}
inline void fb_set_pixel_blended(Fb& fb, uint64_t x, uint64_t y, Col c, uint32_t mode) {   // render.rs:312-333
    if (x < fb.width && y < fb.height) {
        uint64_t idx = (y * fb.width + x) * 4;
        fb_put(fb, idx, col_blend(c, fb_back(fb, idx), mode));
    }
}
inline void fb_set_pixel_alpha(Fb& fb, uint64_t x, uint64_t y, Col color, uint8_t alpha) {   // render.rs:646-667
    if (x < fb.width && y < fb.height) {
        uint64_t idx = (y * fb.width + x) * 4;
        uint8_t back_r = fb.pixels[idx], back_g = fb.pixels[idx + 1], back_b = fb.pixels[idx + 2];
        uint16_t a = alpha, inv_a = (uint16_t)(255 - a);
        fb.pixels[idx]     = (uint8_t)(((uint16_t)color.r * a + (uint16_t)back_r * inv_a) / 255);
        fb.pixels[idx + 1] = (uint8_t)(((uint16_t)color.g * a + (uint16_t)back_g * inv_a) / 255);
        fb.pixels[idx + 2] = (uint8_t)(((uint16_t)color.b * a + (uint16_t)back_b * inv_a) / 255);
        fb.pixels[idx + 3] = 255;
    }
}
// The overlay line family, each restated on its own: draw_line_alpha (render.rs:684-711), draw_line_blended
// (:719-751), draw_line_3d_impl (:767-817), draw_line_3d_alpha (:822-872)
void fb_draw_line_alpha(Fb& fb, int32_t x0, int32_t y0, int32_t x1, int32_t y1, Col color, uint8_t alpha) {
    int32_t dx = std::abs(x1 - x0), dy = -std::abs(y1 - y0);
    int32_t sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
    int32_t err = dx + dy, x = x0, y = y0;
    for (;;) {
        if (x >= 0 && x < (int32_t)fb.width && y >= 0 && y < (int32_t)fb.height) fb_set_pixel_alpha(fb, (uint64_t)x, (uint64_t)y, color, alpha);
        if (x == x1 && y == y1) break;
        int32_t e2 = 2 * err;
        if (e2 >= dy) { err += dy; x += sx; }
        if (e2 <= dx) { err += dx; y += sy; }
    }
}
void fb_draw_line_blended(Fb& fb, int32_t x0, int32_t y0, int32_t x1, int32_t y1, Col color, uint32_t mode) {
    int32_t dx = std::abs(x1 - x0), dy = -std::abs(y1 - y0);
    int32_t sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
    int32_t err = dx + dy, x = x0, y = y0;
    for (;;) {
        if (x >= 0 && x < (int32_t)fb.width && y >= 0 && y < (int32_t)fb.height) {
            if (mode == B32_BLEND_OPAQUE) fb_set_pixel(fb, (uint64_t)x, (uint64_t)y, color);
            else fb_set_pixel_blended(fb, (uint64_t)x, (uint64_t)y, color, mode);
        }
        if (x == x1 && y == y1) break;
        int32_t e2 = 2 * err;
        if (e2 >= dy) { err += dy; x += sx; }
        if (e2 <= dx) { err += dx; y += sy; }
    }
}
void fb_draw_line_3d_impl(Fb& fb, int32_t x0, int32_t y0, float z0, int32_t x1, int32_t y1, float z1, Col color, bool allow_equal) {
    int32_t dx = std::abs(x1 - x0), dy = -std::abs(y1 - y0);
    int32_t sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
    int32_t err = dx + dy, x = x0, y = y0;
    float total_steps = (float)std::max(dx, std::max(-dy, 1));
    float step = 0.0f;
    for (;;) {
        if (x >= 0 && x < (int32_t)fb.width && y >= 0 && y < (int32_t)fb.height) {
            float t = step / total_steps;
            float z = z0 + t * (z1 - z0);
            uint64_t idx = (uint64_t)y * fb.width + (uint64_t)x;
            bool passes = allow_equal ? z <= fb.zbuffer[idx] : z < fb.zbuffer[idx];
            if (passes) fb_set_pixel(fb, (uint64_t)x, (uint64_t)y, color);
        }
        if (x == x1 && y == y1) break;
        int32_t e2 = 2 * err;
        if (e2 >= dy) { err += dy; x += sx; step += 1.0f; }
        if (e2 <= dx) { err += dx; y += sy; if (e2 < dy) step += 1.0f; }
    }
}
void fb_draw_line_3d_alpha(Fb& fb, int32_t x0, int32_t y0, float z0, int32_t x1, int32_t y1, float z1, Col color, uint8_t alpha) {
    const float DEPTH_BIAS = 0.995f;
    float z0_biased = z0 * DEPTH_BIAS, z1_biased = z1 * DEPTH_BIAS;
    int32_t dx = std::abs(x1 - x0), dy = -std::abs(y1 - y0);
    int32_t sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
    int32_t err = dx + dy, x = x0, y = y0;
    float total_steps = (float)std::max(dx, std::max(-dy, 1));
    float step = 0.0f;
    for (;;) {
        if (x >= 0 && x < (int32_t)fb.width && y >= 0 && y < (int32_t)fb.height) {
            float t = step / total_steps;
            float z = z0_biased + t * (z1_biased - z0_biased);
            uint64_t idx = (uint64_t)y * fb.width + (uint64_t)x;
            if (z <= fb.zbuffer[idx]) fb_set_pixel_alpha(fb, (uint64_t)x, (uint64_t)y, color, alpha);
        }
        if (x == x1 && y == y1) break;
        int32_t e2 = 2 * err;
        if (e2 >= dy) { err += dy; x += sx; step += 1.0f; }
        if (e2 <= dx) { err += dx; y += sy; if (e2 < dy) step += 1.0f; }
    }
}

// The filled primitives of the overlay family: draw_circle (render.rs:631-644), draw_circle_alpha (:670-682),
// draw_thick_line (:875-938), draw_filled_rect (:954-972)
void fb_draw_circle(Fb& fb, int32_t cx, int32_t cy, int32_t radius, Col color) {
    int32_t r_sq = radius * radius;
    for (int32_t y = std::max(cy - radius, 0); y <= std::min(cy + radius, (int32_t)fb.height - 1); ++y)
        for (int32_t x = std::max(cx - radius, 0); x <= std::min(cx + radius, (int32_t)fb.width - 1); ++x) {
            int32_t dx = x - cx, dy = y - cy;
            if (dx * dx + dy * dy <= r_sq) fb_set_pixel(fb, (uint64_t)x, (uint64_t)y, color);
        }
}
void fb_draw_circle_alpha(Fb& fb, int32_t cx, int32_t cy, int32_t radius, Col color, uint8_t alpha) {
    int32_t r_sq = radius * radius;
    for (int32_t y = std::max(cy - radius, 0); y <= std::min(cy + radius, (int32_t)fb.height - 1); ++y)
        for (int32_t x = std::max(cx - radius, 0); x <= std::min(cx + radius, (int32_t)fb.width - 1); ++x) {
            int32_t dx = x - cx, dy = y - cy;
            if (dx * dx + dy * dy <= r_sq) fb_set_pixel_alpha(fb, (uint64_t)x, (uint64_t)y, color, alpha);
        }
}
void fb_draw_filled_rect(Fb& fb, int32_t x0, int32_t y0, int32_t x1, int32_t y1, Col color) {
    int32_t min_x = x0 < x1 ? x0 : x1, max_x = x0 < x1 ? x1 : x0;
    int32_t min_y = y0 < y1 ? y0 : y1, max_y = y0 < y1 ? y1 : y0;
    min_x = std::max(min_x, 0); min_y = std::max(min_y, 0);
    max_x = std::min(max_x, (int32_t)fb.width - 1); max_y = std::min(max_y, (int32_t)fb.height - 1);
    for (int32_t y = min_y; y <= max_y; ++y)
        for (int32_t x = min_x; x <= max_x; ++x) fb_set_pixel(fb, (uint64_t)x, (uint64_t)y, color);
}
void fb_draw_thick_line(Fb& fb, int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t thickness, Col color) {
    if (thickness <= 1) { fb_draw_line_blended(fb, x0, y0, x1, y1, color, B32_BLEND_OPAQUE); return; }     // draw_line (:876-879)
    float dx = (float)(x1 - x0), dy = (float)(y1 - y0);
    float len = std::sqrt(dx * dx + dy * dy);
    if (len < 0.001f) return;
    float half = (float)thickness * 0.5f;
    float px = -dy / len * half, py = dx / len * half;
    float c[4][2] = {{(float)x0 + px, (float)y0 + py}, {(float)x0 - px, (float)y0 - py}, {(float)x1 - px, (float)y1 - py}, {(float)x1 + px, (float)y1 + py}};
    float fminx = INFINITY, fmaxx = -INFINITY, fminy = INFINITY, fmaxy = -INFINITY;
    for (auto& k : c) { fminx = rmin(fminx, k[0]); fmaxx = rmax(fmaxx, k[0]); fminy = rmin(fminy, k[1]); fmaxy = rmax(fmaxy, k[1]); }
    int32_t min_x = std::max(f2i32(fminx), 0), max_x = std::min(f2i32(fmaxx), (int32_t)fb.width - 1);
    int32_t min_y = std::max(f2i32(fminy), 0), max_y = std::min(f2i32(fmaxy), (int32_t)fb.height - 1);
    if (min_x > max_x || min_y > max_y) return;
    for (int32_t yy = min_y; yy <= max_y; ++yy)
        for (int32_t xx = min_x; xx <= max_x; ++xx) {
            float p0 = (float)xx + 0.5f, p1 = (float)yy + 0.5f;
            bool inside = true;
            for (int i = 0; i < 4; ++i) {
                const float* a = c[i]; const float* b = c[(i + 1) % 4];
                float cross = (b[0] - a[0]) * (p1 - a[1]) - (b[1] - a[1]) * (p0 - a[0]);
                if (cross < 0.0f) { inside = false; break; }
            }
            if (inside) fb_set_pixel(fb, (uint64_t)xx, (uint64_t)yy, color);
        }
}

// shared tail of the two editor-alpha writers (render.rs:349-373 / 395-419): PS1 blend, then a float lerp
inline void editor_alpha_write8(Fb& fb, uint64_t idx, Col c, uint32_t mode, uint8_t editor_alpha) {
    Col back = fb_back(fb, idx);
    Col ps1 = col_blend(c, back, mode);
    Col fin = ps1;
    if (editor_alpha < 255) {
        float a = (float)editor_alpha / 255.0f;
        float inv_a = 1.0f - a;
        fin = Col{f2u8((float)ps1.r * a + (float)back.r * inv_a), f2u8((float)ps1.g * a + (float)back.g * inv_a),
                  f2u8((float)ps1.b * a + (float)back.b * inv_a), B32_BLEND_OPAQUE};
    }
    fb_put(fb, idx, fin);
}
inline void fb_set_pixel_with_editor_alpha(Fb& fb, uint64_t x, uint64_t y, Col c, uint32_t mode, uint8_t ea) {   // render.rs:338-380
    if (ea == 0) return;
    if (x >= fb.width || y >= fb.height) return;
    editor_alpha_write8(fb, (y * fb.width + x) * 4, c, mode, ea);
}
inline bool fb_set_pixel_with_depth_and_editor_alpha(Fb& fb, uint64_t x, uint64_t y, float z, Col c, uint32_t mode, uint8_t ea) {  // :383-420
    if (ea == 0) return false;
    if (x >= fb.width || y >= fb.height) return false;
    uint64_t depth_idx = y * fb.width + x;
    if (z >= fb.zbuffer[depth_idx]) return false;
    fb.zbuffer[depth_idx] = z;
    editor_alpha_write8(fb, depth_idx * 4, c, mode, ea);
    return true;
}
inline bool fb_set_pixel_with_depth(Fb& fb, uint64_t x, uint64_t y, float z, Col c) {      // render.rs:422-437
    if (x < fb.width && y < fb.height) {
        uint64_t idx = y * fb.width + x;
        if (z < fb.zbuffer[idx]) {
            fb.zbuffer[idx] = z;
            fb_put(fb, idx * 4, c);
            return true;
        }
    }
    return false;
}

// render.rs:1202-1433
void rasterize_triangle(Fb& fb, const Surface& surface, const Tex8* texture, const b32_settings& settings) {
    uint64_t min_x = f2usize(rmax(rmin(rmin(surface.v1.x, surface.v2.x), surface.v3.x), 0.0f));            // :1209-1212
    uint64_t max_x = f2usize(rmin(rmax(rmax(surface.v1.x, surface.v2.x), surface.v3.x) + 1.0f, (float)fb.width));
    uint64_t min_y = f2usize(rmax(rmin(rmin(surface.v1.y, surface.v2.y), surface.v3.y), 0.0f));
    uint64_t max_y = f2usize(rmin(rmax(rmax(surface.v1.y, surface.v2.y), surface.v3.y) + 1.0f, (float)fb.height));
    if (min_x >= max_x || min_y >= max_y) return;                                              // :1215-1217

    float flat_shade[3] = {1.0f, 1.0f, 1.0f};
    if (settings.shading == B32_SHADE_FLAT) {                                                  // :1220-1226
        V3 center_pos = scale(add(add(surface.w1, surface.w2), surface.w3), 1.0f / 3.0f);
        V3 world_normal = normalize(scale(add(add(surface.wn1, surface.wn2), surface.wn3), 1.0f / 3.0f));
        shade_multi_light_color(world_normal, center_pos, settings.lights, settings.n_lights, settings.ambient, flat_shade);
    }
    float gs1[3], gs2[3], gs3[3];
    bool gouraud = settings.shading == B32_SHADE_GOURAUD;
    if (gouraud) {                                                                             // :1229-1237
        shade_multi_light_color(surface.wn1, surface.w1, settings.lights, settings.n_lights, settings.ambient, gs1);
        shade_multi_light_color(surface.wn2, surface.w2, settings.lights, settings.n_lights, settings.ambient, gs2);
        shade_multi_light_color(surface.wn3, surface.w3, settings.lights, settings.n_lights, settings.ambient, gs3);
    }
    bool needs_dither = settings.dithering && (gouraud || texture != nullptr ||               // :1241-1246
                                               !col_eq(surface.vc1, surface.vc2) || !col_eq(surface.vc2, surface.vc3));

    V3 v1 = surface.v1, v2 = surface.v2, v3 = surface.v3;
    float area = (v2.y - v3.y) * (v1.x - v3.x) + (v3.x - v2.x) * (v1.y - v3.y);              // :1257
    if (std::fabs(area) < 0.00001f) return;
    float inv_area = 1.0f / area;
    float a0 = v2.y - v3.y, b0 = v3.x - v2.x, a1 = v3.y - v1.y, b1 = v1.x - v3.x;            // :1265-1270
    float start_x = (float)min_x, start_y = (float)min_y;
    float w0_row = a0 * (start_x - v3.x) + b0 * (start_y - v3.y);                             // :1278-1279
    float w1_row = a1 * (start_x - v3.x) + b1 * (start_y - v3.y);

    for (uint64_t y = min_y; y < max_y; ++y) {                                                 // :1291
        float w0 = w0_row, w1 = w1_row;
        for (uint64_t x = min_x; x < max_x; ++x) {
            float bc_x = w0 * inv_area;
            float bc_y = w1 * inv_area;
            float bc_z = 1.0f - bc_x - bc_y;
            const float ERR = -0.0001f;
            if (bc_x >= ERR && bc_y >= ERR && bc_z >= ERR) {                                   // :1303
                float inv_z1 = 1.0f / v1.z, inv_z2 = 1.0f / v2.z, inv_z3 = 1.0f / v3.z;
                float inv_z_interp = bc_x * inv_z1 + bc_y * inv_z2 + bc_z * inv_z3;
                float z = 1.0f / inv_z_interp;

                if (settings.use_zbuffer && !settings.xray_mode) {                             // :1313-1320
                    uint64_t idx = y * fb.width + x;
                    if (z >= fb.zbuffer[idx]) { w0 += a0; w1 += a1; continue; }
                }

                float u, v;
                if (settings.affine_textures) {                                                // :1323-1340
                    u = bc_x * surface.uv1[0] + bc_y * surface.uv2[0] + bc_z * surface.uv3[0];
                    v = bc_x * surface.uv1[1] + bc_y * surface.uv2[1] + bc_z * surface.uv3[1];
                } else {
                    float u_over_z = bc_x * surface.uv1[0] * inv_z1 + bc_y * surface.uv2[0] * inv_z2 + bc_z * surface.uv3[0] * inv_z3;
                    float v_over_z = bc_x * surface.uv1[1] * inv_z1 + bc_y * surface.uv2[1] * inv_z2 + bc_z * surface.uv3[1] * inv_z3;
                    u = u_over_z / inv_z_interp;
                    v = v_over_z / inv_z_interp;
                }

                Col color = texture ? tex8_sample(*texture, u, 1.0f - v) : Col{255, 255, 255, B32_BLEND_OPAQUE};   // :1343-1347
                if (color.blend == B32_BLEND_ERASE) { w0 += a0; w1 += a1; continue; }          // :1350-1354

                Col vertex_color{                                                              // :1357-1362
                    f2u8(bc_x * (float)surface.vc1.r + bc_y * (float)surface.vc2.r + bc_z * (float)surface.vc3.r),
                    f2u8(bc_x * (float)surface.vc1.g + bc_y * (float)surface.vc2.g + bc_z * (float)surface.vc3.g),
                    f2u8(bc_x * (float)surface.vc1.b + bc_y * (float)surface.vc2.b + bc_z * (float)surface.vc3.b),
                    B32_BLEND_OPAQUE};
                color = col_modulate(color, vertex_color);                                     // :1365

                float shade_r, shade_g, shade_b;                                               // :1368-1381
                if (settings.shading == B32_SHADE_NONE) { shade_r = shade_g = shade_b = 1.0f; }
                else if (settings.shading == B32_SHADE_FLAT) { shade_r = flat_shade[0]; shade_g = flat_shade[1]; shade_b = flat_shade[2]; }
                else {
                    shade_r = bc_x * gs1[0] + bc_y * gs2[0] + bc_z * gs3[0];
                    shade_g = bc_x * gs1[1] + bc_y * gs2[1] + bc_z * gs3[1];
                    shade_b = bc_x * gs1[2] + bc_y * gs2[2] + bc_z * gs3[2];
                }
                color = shade_color_rgb(color, shade_r, shade_g, shade_b);                     // :1383
                if (needs_dither) color = apply_dither(color, x, y);                           // :1387-1389

                uint8_t editor_alpha = surface.editor_alpha;                                   // :1392-1398
                if (editor_alpha == 0) { w0 += a0; w1 += a1; continue; }

                if (settings.use_zbuffer) {                                                    // :1400-1413
                    if (editor_alpha < 255) {
                        fb_set_pixel_with_depth_and_editor_alpha(fb, x, y, z, color, color.blend, editor_alpha);
                    } else if (color.blend == B32_BLEND_OPAQUE) {
                        fb_set_pixel_with_depth(fb, x, y, z, color);
                    } else {
                        uint64_t idx = y * fb.width + x;
                        if (z < fb.zbuffer[idx]) {
                            fb.zbuffer[idx] = z;
                            fb_set_pixel_blended(fb, x, y, color, color.blend);
                        }
                    }
                } else {                                                                       // :1414-1423
                    if (editor_alpha < 255) fb_set_pixel_with_editor_alpha(fb, x, y, color, color.blend, editor_alpha);
                    else if (color.blend == B32_BLEND_OPAQUE) fb_set_pixel(fb, x, y, color);
                    else fb_set_pixel_blended(fb, x, y, color, color.blend);
                }
            }
            w0 += a0; w1 += a1;                                                                // :1427-1428
        }
        w0_row += b0; w1_row += b1;                                                            // :1432-1433
    }
}

// render.rs:2266-2275
inline float calculate_fog_factor(float z, float fog_start, float fog_falloff) {
    if (z <= fog_start) return 0.0f;
    else if (fog_falloff <= 0.0f) return 1.0f;
    else return rmin((z - fog_start) / fog_falloff, 1.0f);
}
// render.rs:2279-2293
inline Col apply_fog_to_color(Col color, Col fog_color, float f) {
    if (f <= 0.0f) return color;
    if (f >= 1.0f) return fog_color;
    float inv = 1.0f - f;
    uint8_t r = f2u8((float)color.r * inv + (float)fog_color.r * f);
    uint8_t g = f2u8((float)color.g * inv + (float)fog_color.g * f);
    uint8_t b = f2u8((float)color.b * inv + (float)fog_color.b * f);
    return Col{r, g, b, B32_BLEND_OPAQUE};
}

inline double now_s() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

struct Projected { std::vector<V3> cam_pos, cam_normals, projected; };

// render.rs:2316-2362
void transform_phase(const b32_vertex* vertices, uint32_t nv, const b32_camera* camera, const b32_settings* settings,
                     uint32_t fbw, uint32_t fbh, Projected& out) {
    out.cam_pos.clear(); out.cam_normals.clear(); out.projected.clear();
    out.cam_pos.reserve(nv); out.cam_normals.reserve(nv); out.projected.reserve(nv);
    V3 cpos = mk3(camera->position), bx = mk3(camera->basis_x), by = mk3(camera->basis_y), bz = mk3(camera->basis_z);
    for (uint32_t i = 0; i < nv; ++i) {
        const b32_vertex& v = vertices[i];
        V3 pos = mk3(v.pos);
        V3 screen_pos, cam_pos;
        if (settings->ortho_enabled) {
            V3 rel = sub(pos, cpos);
            cam_pos = perspective_transform(rel, bx, by, bz);
            screen_pos = project_ortho(cam_pos, settings->ortho_zoom, settings->ortho_center_x, settings->ortho_center_y, fbw, fbh);
        } else if (settings->use_fixed_point) {
            int32_t sx, sy; float fixed_depth;
            project_fixed(pos, camera, fbw, fbh, &sx, &sy, &fixed_depth);
            V3 rel = sub(pos, cpos);
            cam_pos = perspective_transform(rel, bx, by, bz);
            screen_pos = V3{(float)sx, (float)sy, cam_pos.z + 5.0f};
        } else {
            V3 rel = sub(pos, cpos);
            cam_pos = perspective_transform(rel, bx, by, bz);
            screen_pos = project(cam_pos, fbw, fbh);
        }
        out.cam_pos.push_back(cam_pos);
        out.projected.push_back(screen_pos);
        V3 cam_normal = perspective_transform(mk3(v.normal), bx, by, bz);
        out.cam_normals.push_back(normalize(cam_normal));
    }
}

struct Tri3 { V3 a, b, c; };
// WIREFRAME PHASE, identical in render_mesh_15 (render.rs:2574-2635) and render_mesh (render.rs:2195-2256)
void wireframe_phase(Fb& fb, const b32_settings& st, const std::vector<Tri3>& backface_wireframes, const std::vector<Tri3>& frontface_wireframes) {
    const b32_settings* settings = &st;
    struct Edge { int32_t x0, y0; float z0; int32_t x1, y1; float z1; };
    auto collect = [](const std::vector<Tri3>& tris, std::vector<Edge>& unique_edges) {
        for (const Tri3& t : tris) {
            Edge es[3] = {{f2i32(t.a.x), f2i32(t.a.y), t.a.z, f2i32(t.b.x), f2i32(t.b.y), t.b.z},
                          {f2i32(t.b.x), f2i32(t.b.y), t.b.z, f2i32(t.c.x), f2i32(t.c.y), t.c.z},
                          {f2i32(t.c.x), f2i32(t.c.y), t.c.z, f2i32(t.a.x), f2i32(t.a.y), t.a.z}};
            for (const Edge& e0 : es) {
                bool lt = (e0.x0 < e0.x1) || (e0.x0 == e0.x1 && e0.y0 < e0.y1);      // tuple `<`
                Edge e = lt ? e0 : Edge{e0.x1, e0.y1, e0.z1, e0.x0, e0.y0, e0.z0};
                bool found = false;
                for (const Edge& u : unique_edges) if (u.x0 == e.x0 && u.y0 == e.y0 && u.x1 == e.x1 && u.y1 == e.y1) { found = true; break; }
                if (!found) unique_edges.push_back(e);
            }
        }
    };
    if (settings->backface_cull && settings->backface_wireframe) {
        std::vector<Edge> ue; collect(backface_wireframes, ue);
        Col wc{80, 80, 100, B32_BLEND_OPAQUE};
        for (const Edge& e : ue) fb_draw_line_3d(fb, e.x0, e.y0, e.z0, e.x1, e.y1, e.z1, wc);
    }
    if (settings->wireframe_overlay && !frontface_wireframes.empty()) {
        std::vector<Edge> ue; collect(frontface_wireframes, ue);
        Col wc{200, 200, 220, B32_BLEND_OPAQUE};
        for (const Edge& e : ue) fb_draw_line(fb, e.x0, e.y0, e.x1, e.y1, wc);
    }
}

inline Col vcol(const b32_vertex& v) { return Col{v.r, v.g, v.b, v.blend}; }

}  // namespace

// =============================================================================================
// exported
// =============================================================================================
extern "C" {

void b32o_set_compat(uint32_t flags) { g_compat = flags; }
uint32_t b32o_get_compat(void) { return g_compat; }
uint8_t b32o_unr_table(uint32_t i) { return UNR.t[i < 257 ? i : 256]; }
int32_t b32o_fixed_from_f32(float f) { return fx_from_f32(f); }
int32_t b32o_fixed_mul(int32_t a, int32_t b) { return fx_mul(a, b); }
int32_t b32o_div_unr(int32_t num, int32_t den) { return fx_div_unr(num, den); }
void b32o_project_fixed(const float world[3], const b32_camera* cam, uint32_t w, uint32_t h, int32_t* sx, int32_t* sy, float* depth) {
    project_fixed(mk3(world), cam, w, h, sx, sy, depth);
}
void b32o_dither_and_quantize(uint8_t r8, uint8_t g8, uint8_t b8, uint32_t x, uint32_t y, uint8_t out5[3]) {
    dither_and_quantize(r8, g8, b8, x, y, out5);
}
void b32o_blend_rgb555(uint8_t fr, uint8_t fg, uint8_t fb, uint8_t br, uint8_t bg, uint8_t bb, uint32_t mode, uint8_t out8[3]) {
    blend_rgb555(fr, fg, fb, br, bg, bb, mode, out8);
}
uint16_t b32o_texture_sample(const b32_tex_desc* tex, float u, float v) {
    Tex15 t = to_texture15(*tex);
    return tex_sample(t, u, v);
}
void b32o_shade_multi_light(const float normal[3], const float world_pos[3], const b32_light* lights, uint32_t n_lights,
                            float ambient, float out_rgb[3]) {
    shade_multi_light_color(mk3(normal), mk3(world_pos), lights, n_lights, ambient, out_rgb);
}

void b32o_acosf(const float* x, float* out, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) out[i] = ref_acosf(x[i]);
}

void b32o_fb_clear(uint8_t* rgba, float* z, uint32_t w, uint32_t h, uint8_t r, uint8_t g, uint8_t b, uint8_t a) {
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        rgba[i * 4] = r; rgba[i * 4 + 1] = g; rgba[i * 4 + 2] = b; rgba[i * 4 + 3] = a;
        if (z) z[i] = std::numeric_limits<float>::max();
    }
}

void b32o_transform(const b32_vertex* v, uint32_t nv, const b32_camera* cam, const b32_settings* s, uint32_t w, uint32_t h,
                    float* out_screen, float* out_cam) {
    Projected p;
    transform_phase(v, nv, cam, s, w, h, p);
    for (uint32_t i = 0; i < nv; ++i) {
        out_screen[i * 3] = p.projected[i].x; out_screen[i * 3 + 1] = p.projected[i].y; out_screen[i * 3 + 2] = p.projected[i].z;
        out_cam[i * 3] = p.cam_pos[i].x; out_cam[i * 3 + 1] = p.cam_pos[i].y; out_cam[i * 3 + 2] = p.cam_pos[i].z;
    }
}

int b32o_render_mesh_15(uint8_t* fb_rgba, float* fb_z, uint32_t w, uint32_t h,
                        const b32_vertex* vertices, uint32_t nv, const b32_face* faces, uint32_t nf,
                        const b32_tex_desc* textures, uint32_t ntex, const b32_camera* camera,
                        const b32_settings* settings, const b32_fog* fog, b32_timings* timings,
                        uint32_t* draw_order, uint32_t cap, uint32_t* n_drawn) {
    Fb fb{fb_rgba, fb_z, w, h};
    b32_timings tm{};

    // `textures: &[Texture15]` — indexed inputs are expanded through their CLUT first, exactly as the
    // callers do before render_mesh_15 (src/scene.rs:161-165).  Not part of the timed phases.
    std::vector<Tex15> tex15;
    tex15.reserve(ntex);
    for (uint32_t i = 0; i < ntex; ++i) tex15.push_back(to_texture15(textures[i]));

    // === TRANSFORM PHASE === render.rs:2313-2362
    double t0 = now_s();
    Projected P;
    transform_phase(vertices, nv, camera, settings, w, h, P);
    tm.transform_ms = (float)((now_s() - t0) * 1000.0);

    // === CULL PHASE === render.rs:2364-2516
    double cull_start = now_s();
    double fog_total = 0.0;
    std::vector<Surface> surfaces;
    surfaces.reserve(nf);
    std::vector<Tri3> backface_wireframes, frontface_wireframes;

    for (uint32_t face_idx = 0; face_idx < nf; ++face_idx) {
        const b32_face& face = faces[face_idx];
        if (face.v0 >= nv || face.v1 >= nv || face.v2 >= nv) return B32_ERR_OOB_INDEX;  // Rust: index panic
        uint32_t tex_id = face.flags & 0xFFFFu;
        uint32_t face_blend = (face.flags >> 16) & 7u;
        bool black_transparent = ((face.flags >> 19) & 1u) != 0;
        uint8_t editor_alpha = (uint8_t)(face.flags >> 24);
        const Tex15* tex = (tex_id != B32_FACE_TEX_NONE && tex_id < ntex) ? &tex15[tex_id] : nullptr;

        V3 cv1 = P.cam_pos[face.v0], cv2 = P.cam_pos[face.v1], cv3 = P.cam_pos[face.v2];
        if (!settings->ortho_enabled) {                                                  // :2380-2385
            if (cv1.z <= NEAR_PLANE || cv2.z <= NEAR_PLANE || cv3.z <= NEAR_PLANE) continue;
        }
        V3 v1 = P.projected[face.v0], v2 = P.projected[face.v1], v3 = P.projected[face.v2];
        float signed_area = (v2.x - v1.x) * (v3.y - v1.y) - (v3.x - v1.x) * (v2.y - v1.y);  // :2393
        bool is_backface = signed_area <= 0.0f;

        V3 edge1 = sub(cv2, cv1), edge2 = sub(cv3, cv1);                                 // :2397-2399 (dead value)
        V3 cr{edge1.y * edge2.z - edge1.z * edge2.y, edge1.z * edge2.x - edge1.x * edge2.z, edge1.x * edge2.y - edge1.y * edge2.x};
        V3 normal = normalize(cr);

        bool has_transparency;                                                           // :2403-2415
        if (tex && tex->blend_mode != B32_BLEND_OPAQUE) has_transparency = true;
        else if (face_blend != B32_BLEND_OPAQUE) has_transparency = true;
        else has_transparency = editor_alpha < 255;
        if (g_compat & COMPAT_TRANSP_TEX_ONLY)
            has_transparency = tex ? tex->blend_mode != B32_BLEND_OPAQUE : face_blend != B32_BLEND_OPAQUE;

        double fog_t0 = now_s();
        Col vc1, vc2, vc3;                                                               // :2419-2443
        if (fog) {
            if (cv1.z > fog->cull_distance && cv2.z > fog->cull_distance && cv3.z > fog->cull_distance) {
                fog_total += now_s() - fog_t0;
                continue;
            }
            float f1 = calculate_fog_factor(cv1.z, fog->start, fog->falloff);
            float f2 = calculate_fog_factor(cv2.z, fog->start, fog->falloff);
            float f3 = calculate_fog_factor(cv3.z, fog->start, fog->falloff);
            Col fc{fog->r, fog->g, fog->b, fog->blend};
            vc1 = apply_fog_to_color(vcol(vertices[face.v0]), fc, f1);
            vc2 = apply_fog_to_color(vcol(vertices[face.v1]), fc, f2);
            vc3 = apply_fog_to_color(vcol(vertices[face.v2]), fc, f3);
        } else {
            vc1 = vcol(vertices[face.v0]); vc2 = vcol(vertices[face.v1]); vc3 = vcol(vertices[face.v2]);
        }
        fog_total += now_s() - fog_t0;

        const b32_vertex &A = vertices[face.v0], &B = vertices[face.v1], &C = vertices[face.v2];
        if (is_backface) {                                                               // :2445-2481
            if (!settings->xray_mode) backface_wireframes.push_back(Tri3{v1, v2, v3});
            if (!settings->backface_cull || settings->xray_mode) {
                Surface s;
                s.v1 = v1; s.v2 = v3; s.v3 = v2;
                s.w1 = mk3(A.pos); s.w2 = mk3(C.pos); s.w3 = mk3(B.pos);
                s.vn1 = scale(P.cam_normals[face.v0], -1.0f); s.vn2 = scale(P.cam_normals[face.v2], -1.0f); s.vn3 = scale(P.cam_normals[face.v1], -1.0f);
                s.wn1 = scale(mk3(A.normal), -1.0f); s.wn2 = scale(mk3(C.normal), -1.0f); s.wn3 = scale(mk3(B.normal), -1.0f);
                s.uv1[0] = A.uv[0]; s.uv1[1] = A.uv[1]; s.uv2[0] = C.uv[0]; s.uv2[1] = C.uv[1]; s.uv3[0] = B.uv[0]; s.uv3[1] = B.uv[1];
                s.vc1 = vc1; s.vc2 = vc3; s.vc3 = vc2;
                s.normal = scale(normal, -1.0f);
                s.face_idx = face_idx; s.black_transparent = black_transparent; s.has_transparency = has_transparency;
                s.blend_mode = face_blend; s.editor_alpha = editor_alpha;
                surfaces.push_back(s);
            }
        } else {                                                                         // :2482-2512
            Surface s;
            s.v1 = v1; s.v2 = v2; s.v3 = v3;
            s.w1 = mk3(A.pos); s.w2 = mk3(B.pos); s.w3 = mk3(C.pos);
            s.vn1 = P.cam_normals[face.v0]; s.vn2 = P.cam_normals[face.v1]; s.vn3 = P.cam_normals[face.v2];
            s.wn1 = mk3(A.normal); s.wn2 = mk3(B.normal); s.wn3 = mk3(C.normal);
            s.uv1[0] = A.uv[0]; s.uv1[1] = A.uv[1]; s.uv2[0] = B.uv[0]; s.uv2[1] = B.uv[1]; s.uv3[0] = C.uv[0]; s.uv3[1] = C.uv[1];
            s.vc1 = vc1; s.vc2 = vc2; s.vc3 = vc3;
            s.normal = normal;
            s.face_idx = face_idx; s.black_transparent = black_transparent; s.has_transparency = has_transparency;
            s.blend_mode = face_blend; s.editor_alpha = editor_alpha;
            surfaces.push_back(s);
            if (settings->wireframe_overlay) frontface_wireframes.push_back(Tri3{v1, v2, v3});
        }
    }
    tm.cull_ms = (float)((now_s() - cull_start) * 1000.0);
    tm.fog_ms = (float)(fog_total * 1000.0);

    // === SORT PHASE === render.rs:2518-2545
    double sort_start = now_s();
    std::vector<Surface> opaque_surfaces, transparent_surfaces;      // Iterator::partition keeps order
    opaque_surfaces.reserve(surfaces.size());
    for (const Surface& s : surfaces) (s.has_transparency ? transparent_surfaces : opaque_surfaces).push_back(s);
    surfaces.clear(); surfaces.shrink_to_fit();

    auto center_z = [](const Surface& s) { return (s.v1.z + s.v2.z + s.v3.z) / 3.0f; };
    // slice::sort_by is a stable merge sort; `b.partial_cmp(a).unwrap()` panics on NaN as soon as a
    // NaN key is compared, which happens for any slice of length >= 2.
    auto sort_back_to_front = [&](std::vector<Surface>& v) -> bool {
        if (v.size() >= 2) for (const Surface& s : v) { float k = center_z(s); if (k != k) return false; }
        std::stable_sort(v.begin(), v.end(), [&](const Surface& a, const Surface& b) { return center_z(a) > center_z(b); });
        return true;
    };
    if (!sort_back_to_front(transparent_surfaces)) return B32_ERR_NAN_DEPTH;
    if (!settings->use_zbuffer) { if (!sort_back_to_front(opaque_surfaces)) return B32_ERR_NAN_DEPTH; }
    tm.sort_ms = (float)((now_s() - sort_start) * 1000.0);
    tm.triangles_drawn = (uint32_t)(opaque_surfaces.size() + transparent_surfaces.size());

    if (n_drawn) *n_drawn = tm.triangles_drawn;
    if (draw_order) {
        uint32_t k = 0;
        for (const Surface& s : opaque_surfaces) { if (k < cap) draw_order[k] = (uint32_t)s.face_idx; ++k; }
        for (const Surface& s : transparent_surfaces) { if (k < cap) draw_order[k] = (uint32_t)s.face_idx; ++k; }
    }

    // === DRAW PHASE === render.rs:2547-2572
    double draw_start = now_s();
    if (!settings->wireframe_overlay) {
        auto tex_of = [&](const Surface& s) -> const Tex15* {
            uint32_t id = faces[s.face_idx].flags & 0xFFFFu;
            return (id != B32_FACE_TEX_NONE && id < ntex) ? &tex15[id] : nullptr;
        };
        for (const Surface& s : opaque_surfaces) rasterize_triangle_15(fb, s, tex_of(s), s.blend_mode, s.black_transparent, *settings, false);
        for (const Surface& s : transparent_surfaces) rasterize_triangle_15(fb, s, tex_of(s), s.blend_mode, s.black_transparent, *settings, true);
    }
    tm.draw_ms = (float)((now_s() - draw_start) * 1000.0);

    // === WIREFRAME PHASE === render.rs:2574-2635
    double wire_start = now_s();
    wireframe_phase(fb, *settings, backface_wireframes, frontface_wireframes);
    tm.wireframe_ms = (float)((now_s() - wire_start) * 1000.0);

    if (timings) *timings = tm;
    return B32_OK;
}

// render_mesh, render.rs:1971-2259
int b32o_render_mesh(uint8_t* fb_rgba, float* fb_z, uint32_t w, uint32_t h,
                     const b32_vertex* vertices, uint32_t nv, const b32_face* faces, uint32_t nf,
                     const b32_tex8_desc* textures, uint32_t ntex, const b32_camera* camera,
                     const b32_settings* settings, b32_timings* timings,
                     uint32_t* draw_order, uint32_t cap, uint32_t* n_drawn) {
    Fb fb{fb_rgba, fb_z, w, h};
    b32_timings tm{};
    std::vector<Tex8> tex8(ntex);
    for (uint32_t i = 0; i < ntex; ++i)
        tex8[i] = Tex8{textures[i].width, textures[i].height, reinterpret_cast<const Col*>(textures[i].pixels), textures[i].blend_mode};

    // === TRANSFORM PHASE === render.rs:1981-2028 (the same loop as render_mesh_15's)
    double t0 = now_s();
    Projected P;
    transform_phase(vertices, nv, camera, settings, w, h, P);
    tm.transform_ms = (float)((now_s() - t0) * 1000.0);

    // === CULL PHASE === render.rs:2030-2151
    double cull_start = now_s();
    std::vector<Surface> surfaces;
    surfaces.reserve(nf);
    std::vector<Tri3> backface_wireframes, frontface_wireframes;
    for (uint32_t face_idx = 0; face_idx < nf; ++face_idx) {
        const b32_face& face = faces[face_idx];
        if (face.v0 >= nv || face.v1 >= nv || face.v2 >= nv) return B32_ERR_OOB_INDEX;  // Rust: index panic
        uint32_t tex_id = face.flags & 0xFFFFu;
        uint32_t face_blend = (face.flags >> 16) & 7u;
        bool black_transparent = ((face.flags >> 19) & 1u) != 0;
        uint8_t editor_alpha = (uint8_t)(face.flags >> 24);
        const Tex8* tex = (tex_id != B32_FACE_TEX_NONE && tex_id < ntex) ? &tex8[tex_id] : nullptr;

        V3 cv1 = P.cam_pos[face.v0], cv2 = P.cam_pos[face.v1], cv3 = P.cam_pos[face.v2];
        if (!settings->ortho_enabled) {                                                  // :2049-2053
            if (cv1.z <= NEAR_PLANE || cv2.z <= NEAR_PLANE || cv3.z <= NEAR_PLANE) continue;
        }
        V3 v1 = P.projected[face.v0], v2 = P.projected[face.v1], v3 = P.projected[face.v2];
        float signed_area = (v2.x - v1.x) * (v3.y - v1.y) - (v3.x - v1.x) * (v2.y - v1.y);  // :2061
        bool is_backface = signed_area <= 0.0f;

        V3 edge1 = sub(cv2, cv1), edge2 = sub(cv3, cv1);                                 // :2065-2067 (dead value)
        V3 cr{edge1.y * edge2.z - edge1.z * edge2.y, edge1.z * edge2.x - edge1.x * edge2.z, edge1.x * edge2.y - edge1.y * edge2.x};
        V3 normal = normalize(cr);

        bool has_transparency = (tex && tex->blend_mode != B32_BLEND_OPAQUE) || editor_alpha < 255;   // :2070-2075 (never read again)

        const b32_vertex &A = vertices[face.v0], &B = vertices[face.v1], &C = vertices[face.v2];
        if (is_backface) {                                                               // :2077-2113
            if (!settings->xray_mode) backface_wireframes.push_back(Tri3{v1, v2, v3});
            if (!settings->backface_cull || settings->xray_mode) {
                Surface s;
                s.v1 = v1; s.v2 = v3; s.v3 = v2;
                s.w1 = mk3(A.pos); s.w2 = mk3(C.pos); s.w3 = mk3(B.pos);
                s.vn1 = scale(P.cam_normals[face.v0], -1.0f); s.vn2 = scale(P.cam_normals[face.v2], -1.0f); s.vn3 = scale(P.cam_normals[face.v1], -1.0f);
                s.wn1 = scale(mk3(A.normal), -1.0f); s.wn2 = scale(mk3(C.normal), -1.0f); s.wn3 = scale(mk3(B.normal), -1.0f);
                s.uv1[0] = A.uv[0]; s.uv1[1] = A.uv[1]; s.uv2[0] = C.uv[0]; s.uv2[1] = C.uv[1]; s.uv3[0] = B.uv[0]; s.uv3[1] = B.uv[1];
                s.vc1 = vcol(A); s.vc2 = vcol(C); s.vc3 = vcol(B);
                s.normal = scale(normal, -1.0f);
                s.face_idx = face_idx; s.black_transparent = black_transparent; s.has_transparency = has_transparency;
                s.blend_mode = face_blend; s.editor_alpha = editor_alpha;
                surfaces.push_back(s);
            }
        } else {                                                                         // :2114-2148
            Surface s;
            s.v1 = v1; s.v2 = v2; s.v3 = v3;
            s.w1 = mk3(A.pos); s.w2 = mk3(B.pos); s.w3 = mk3(C.pos);
            s.vn1 = P.cam_normals[face.v0]; s.vn2 = P.cam_normals[face.v1]; s.vn3 = P.cam_normals[face.v2];
            s.wn1 = mk3(A.normal); s.wn2 = mk3(B.normal); s.wn3 = mk3(C.normal);
            s.uv1[0] = A.uv[0]; s.uv1[1] = A.uv[1]; s.uv2[0] = B.uv[0]; s.uv2[1] = B.uv[1]; s.uv3[0] = C.uv[0]; s.uv3[1] = C.uv[1];
            s.vc1 = vcol(A); s.vc2 = vcol(B); s.vc3 = vcol(C);
            s.normal = normal;
            s.face_idx = face_idx; s.black_transparent = black_transparent; s.has_transparency = has_transparency;
            s.blend_mode = face_blend; s.editor_alpha = editor_alpha;
            surfaces.push_back(s);
            if (settings->wireframe_overlay) frontface_wireframes.push_back(Tri3{v1, v2, v3});
        }
    }
    tm.cull_ms = (float)((now_s() - cull_start) * 1000.0);

    // === SORT PHASE === render.rs:2153-2168: ONE list, sorted only for the painter's algorithm
    double sort_start = now_s();
    auto center_z = [](const Surface& s) { return (s.v1.z + s.v2.z + s.v3.z) / 3.0f; };
    if (!settings->use_zbuffer) {
        if (surfaces.size() >= 2) for (const Surface& s : surfaces) { float k = center_z(s); if (k != k) return B32_ERR_NAN_DEPTH; }   // unwrap() panic
        std::stable_sort(surfaces.begin(), surfaces.end(), [&](const Surface& a, const Surface& b) { return center_z(a) > center_z(b); });
    }
    tm.sort_ms = (float)((now_s() - sort_start) * 1000.0);
    tm.triangles_drawn = (uint32_t)surfaces.size();
    if (n_drawn) *n_drawn = tm.triangles_drawn;
    if (draw_order) { uint32_t k = 0; for (const Surface& s : surfaces) { if (k < cap) draw_order[k] = (uint32_t)s.face_idx; ++k; } }

    // === DRAW PHASE === render.rs:2170-2186
    double draw_start = now_s();
    if (!settings->wireframe_overlay) {
        for (const Surface& s : surfaces) {
            uint32_t id = faces[s.face_idx].flags & 0xFFFFu;
            const Tex8* tex = (id != B32_FACE_TEX_NONE && id < ntex) ? &tex8[id] : nullptr;
            rasterize_triangle(fb, s, tex, *settings);
        }
    }
    tm.draw_ms = (float)((now_s() - draw_start) * 1000.0);

    // === WIREFRAME PHASE === render.rs:2188-2256
    double wire_start = now_s();
    wireframe_phase(fb, *settings, backface_wireframes, frontface_wireframes);
    tm.wireframe_ms = (float)((now_s() - wire_start) * 1000.0);

    if (timings) *timings = tm;
    return B32_OK;
}

// Framebuffer::render_skybox, step 1 (render.rs:89-139), with rasterize_skybox_triangle (render.rs:242-299) inlined
int b32o_render_skybox_mesh(uint8_t* fb_rgba, uint32_t w, uint32_t h, const b32_sky_vertex* vertices, uint32_t nv,
                            const uint32_t* faces, uint32_t nf, const b32_camera* camera) {
    V3 cpos = mk3(camera->position), bx = mk3(camera->basis_x), by = mk3(camera->basis_y), bz = mk3(camera->basis_z);
    struct P3 { float x, y, z; };
    std::vector<P3> projected;
    projected.reserve(nv);
    const float nan = std::numeric_limits<float>::quiet_NaN();
    for (uint32_t i = 0; i < nv; ++i) {                                              // :96-108
        V3 world_pos = mk3(vertices[i].pos);
        V3 rel_pos = sub(world_pos, cpos);
        V3 cam_space = perspective_transform(rel_pos, bx, by, bz);
        if (cam_space.z <= 0.1f) { projected.push_back(P3{nan, nan, nan}); continue; }
        V3 screen = project(cam_space, w, h);
        projected.push_back(P3{screen.x, screen.y, cam_space.z});
    }
    for (uint32_t fi = 0; fi < nf; ++fi) {                                           // :111-139
        uint32_t i0 = faces[fi * 3], i1 = faces[fi * 3 + 1], i2 = faces[fi * 3 + 2];
        if (i0 >= nv || i1 >= nv || i2 >= nv) return B32_ERR_OOB_INDEX;
        P3 p0 = projected[i0], p1 = projected[i1], p2 = projected[i2];
        if (p0.x != p0.x || p1.x != p1.x || p2.x != p2.x) continue;
        float signed_area = (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
        if (signed_area >= 0.0f) continue;
        const b32_sky_vertex &c0 = vertices[i0], &c1 = vertices[i1], &c2 = vertices[i2];
        // rasterize_skybox_triangle, :242-299
        uint64_t min_x = f2usize(rmax(rmin(rmin(p0.x, p1.x), p2.x), 0.0f));
        uint64_t max_x = f2usize(rmin(rmax(rmax(p0.x, p1.x), p2.x), (float)w - 1.0f));
        uint64_t min_y = f2usize(rmax(rmin(rmin(p0.y, p1.y), p2.y), 0.0f));
        uint64_t max_y = f2usize(rmin(rmax(rmax(p0.y, p1.y), p2.y), (float)h - 1.0f));
        if (min_x > max_x || min_y > max_y) continue;
        float denom = (p1.y - p2.y) * (p0.x - p2.x) + (p2.x - p1.x) * (p0.y - p2.y);
        if (std::fabs(denom) < 0.0001f) continue;
        float inv_denom = 1.0f / denom;
        for (uint64_t y = min_y; y <= max_y; ++y) {
            for (uint64_t x = min_x; x <= max_x; ++x) {
                float px = (float)x + 0.5f, py = (float)y + 0.5f;
                float w0 = ((p1.y - p2.y) * (px - p2.x) + (p2.x - p1.x) * (py - p2.y)) * inv_denom;
                float w1 = ((p2.y - p0.y) * (px - p2.x) + (p0.x - p2.x) * (py - p2.y)) * inv_denom;
                float w2 = 1.0f - w0 - w1;
                if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) {
                    uint64_t idx = (y * w + x) * 4;
                    fb_rgba[idx]     = f2u8((float)c0.r * w0 + (float)c1.r * w1 + (float)c2.r * w2);
                    fb_rgba[idx + 1] = f2u8((float)c0.g * w0 + (float)c1.g * w1 + (float)c2.g * w2);
                    fb_rgba[idx + 2] = f2u8((float)c0.b * w0 + (float)c1.b * w1 + (float)c2.b * w2);
                    fb_rgba[idx + 3] = 255;
                }
            }
        }
    }
    return B32_OK;
}

// render_stars from the star's direction on (render.rs:175-199) + draw_star_diamond (:203-235); the LCG and the libm
// calls that produce `dir` and the brightness-scaled colour stay with the caller (see b32_star in b32_raster.h)
int b32o_render_stars(uint8_t* fb_rgba, uint32_t w, uint32_t h, const b32_star* stars, uint32_t n, const b32_camera* camera, float size) {
    Fb fb{fb_rgba, nullptr, w, h};
    V3 bx{camera->basis_x[0], camera->basis_x[1], camera->basis_x[2]};
    V3 by{camera->basis_y[0], camera->basis_y[1], camera->basis_y[2]};
    V3 bz{camera->basis_z[0], camera->basis_z[1], camera->basis_z[2]};
    auto set_pixel_safe = [&](int32_t x, int32_t y, Col c) {                         // :237-241
        if (x >= 0 && y >= 0 && x < (int32_t)fb.width && y < (int32_t)fb.height) fb_set_pixel(fb, (uint64_t)x, (uint64_t)y, c);
    };
    for (uint32_t i = 0; i < n; ++i) {
        const b32_star& st = stars[i];
        V3 dir{st.dir[0], st.dir[1], st.dir[2]};
        V3 cam_space = perspective_transform(V3{dir.x * 10000.0f, dir.y * 10000.0f, dir.z * 10000.0f}, bx, by, bz);   // :178
        if (!(cam_space.z > 0.1f)) continue;                                         // :180
        V3 screen = project(cam_space, w, h);
        Col color{st.r, st.g, st.b, B32_BLEND_OPAQUE};                               // Color::new
        int32_t cx = f2i32(screen.x), cy = f2i32(screen.y);
        int32_t s = f2i32(rmax(size, 1.0f));                                         // :204
        set_pixel_safe(cx, cy, color);
        if (s >= 2) {
            Col dim{f2u8((float)color.r * 0.7f), f2u8((float)color.g * 0.7f), f2u8((float)color.b * 0.7f), B32_BLEND_OPAQUE};
            set_pixel_safe(cx - 1, cy, dim); set_pixel_safe(cx + 1, cy, dim); set_pixel_safe(cx, cy - 1, dim); set_pixel_safe(cx, cy + 1, dim);
        }
        if (s >= 3) {
            Col faint{f2u8((float)color.r * 0.4f), f2u8((float)color.g * 0.4f), f2u8((float)color.b * 0.4f), B32_BLEND_OPAQUE};
            set_pixel_safe(cx - 2, cy, faint); set_pixel_safe(cx + 2, cy, faint); set_pixel_safe(cx, cy - 2, faint); set_pixel_safe(cx, cy + 2, faint);
        }
    }
    return B32_OK;
}

// render_asset_parts' per-object vertex transform, src/scene.rs:121-160 (cos_f / sin_f = facing.cos() / .sin(), computed
// by the caller): returns 1 when has_transform (:123) and `out` holds the transformed copy, 0 when the part is drawn as is
int b32o_place_vertices(const b32_vertex* in, uint32_t nv, float facing, float cos_f, float sin_f, const float world_pos[3], b32_vertex* out) {
    bool has_transform = std::fabs(facing) > 0.0001f || std::fabs(world_pos[0]) > 0.0001f || std::fabs(world_pos[1]) > 0.0001f ||
                         std::fabs(world_pos[2]) > 0.0001f;
    for (uint32_t i = 0; i < nv; ++i) {
        b32_vertex v = in[i];
        if (has_transform) {
            float rx = v.pos[0] * cos_f - v.pos[2] * sin_f;
            float rz = v.pos[0] * sin_f + v.pos[2] * cos_f;
            b32_vertex o = v;
            o.pos[0] = rx + world_pos[0]; o.pos[1] = v.pos[1] + world_pos[1]; o.pos[2] = rz + world_pos[2];
            o.normal[0] = v.normal[0] * cos_f - v.normal[2] * sin_f;
            o.normal[1] = v.normal[1];
            o.normal[2] = v.normal[0] * sin_f + v.normal[2] * cos_f;
            v = o;
        }
        out[i] = v;
    }
    return has_transform ? 1 : 0;
}

// Framebuffer::clear_gradient, render.rs:60-77; Color::lerp, types.rs:811-820
void b32o_fb_clear_gradient(uint8_t* rgba, float* z, uint32_t w, uint32_t h, const uint8_t top[3], const uint8_t bottom[3], uint8_t a) {
    for (uint64_t y = 0; y < h; ++y) {
        float t = h > 1 ? (float)y / (float)(h - 1) : 0.0f;
        t = rmin(rmax(t, 0.0f), 1.0f);
        float inv_t = 1.0f - t;
        uint8_t r = f2u8((float)top[0] * inv_t + (float)bottom[0] * t);
        uint8_t g = f2u8((float)top[1] * inv_t + (float)bottom[1] * t);
        uint8_t b = f2u8((float)top[2] * inv_t + (float)bottom[2] * t);
        for (uint64_t x = 0; x < w; ++x) {
            uint64_t idx = (y * w + x) * 4;
            rgba[idx] = r; rgba[idx + 1] = g; rgba[idx + 2] = b; rgba[idx + 3] = a;
            z[y * w + x] = 3.40282347e+38f;
        }
    }
}

// A list of overlay lines drawn one after the other: what a caller of Framebuffer::draw_line* does
int b32o_draw_lines(uint8_t* fb_rgba, float* fb_z, uint32_t w, uint32_t h, const b32_line* lines, uint32_t n) {
    Fb fb{fb_rgba, fb_z, w, h};
    for (uint32_t i = 0; i < n; ++i) {
        const b32_line& l = lines[i];
        Col c{l.r, l.g, l.b, l.blend};
        switch (l.kind) {
            case B32_LINE_2D:         fb_draw_line_blended(fb, l.x0, l.y0, l.x1, l.y1, c, l.mode); break;   // draw_line = mode Opaque (:714-716)
            case B32_LINE_2D_ALPHA:   fb_draw_line_alpha(fb, l.x0, l.y0, l.x1, l.y1, c, l.alpha); break;
            case B32_LINE_3D:         fb_draw_line_3d_impl(fb, l.x0, l.y0, l.z0, l.x1, l.y1, l.z1, c, false); break;
            case B32_LINE_3D_OVERLAY: fb_draw_line_3d_impl(fb, l.x0, l.y0, l.z0, l.x1, l.y1, l.z1, c, true); break;
            case B32_LINE_3D_ALPHA:   fb_draw_line_3d_alpha(fb, l.x0, l.y0, l.z0, l.x1, l.y1, l.z1, c, l.alpha); break;
            case B32_LINE_CIRCLE:       fb_draw_circle(fb, l.x0, l.y0, l.x1, c); break;
            case B32_LINE_CIRCLE_ALPHA: fb_draw_circle_alpha(fb, l.x0, l.y0, l.x1, c, l.alpha); break;
            case B32_LINE_FILLED_RECT:  fb_draw_filled_rect(fb, l.x0, l.y0, l.x1, l.y1, c); break;
            case B32_LINE_THICK:        fb_draw_thick_line(fb, l.x0, l.y0, l.x1, l.y1, f2i32(l.z0), c); break;
            default: return B32_ERR_INVALID;
        }
    }
    return B32_OK;
}

}  // extern "C"
