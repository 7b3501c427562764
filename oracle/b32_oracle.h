/*
 * b32_oracle.h — CPU oracle for the BONNIE-32 rasterizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (bonnie-32_b200/, include/) may link,
 * import or call this.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use
 * it, and there only as the checker / the timed CPU arm.
 *
 * It is a line-by-line C++ restatement of /root/reference/src/rasterizer (Rust); every function
 * cites the reference lines it follows.  The Rust reference cannot be compiled in this environment
 * (no rustc/cargo; un-vendored crates), and the reference has NO test that pins render.rs, so the
 * fill path is PARITY-UNPINNED by reference-executed vectors.  It is pinned instead by (1) the
 * exact facts of the reference's own unit tests in fixed.rs:477-548, (2) spec constants (UNR table
 * formula, dither matrix), and (3) bit-for-bit agreement with a second, independently written
 * restatement (oracle/pymodel.py, numpy, vectorised per triangle) on golden scenes in tests/golden.
 */
#ifndef B32_ORACLE_H
#define B32_ORACLE_H

#include "../include/b32_raster.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Compatibility switches (bit mask, default 0 = the 0.1.11 source): turn back, one by one, the four source
 * changes that separate /root/reference/docs/bonnie-32.wasm (0.1.8) from the source tree, so the restatement
 * can be compared bit for bit with the reference's own compiled code.  See b32_oracle.cpp COMPAT_*. */
void     b32o_set_compat(uint32_t flags);
uint32_t b32o_get_compat(void);

/* fixed.rs */
uint8_t b32o_unr_table(uint32_t i);                               /* fixed.rs:20-31   */
int32_t b32o_fixed_from_f32(float f);                             /* fixed.rs:125-127 */
int32_t b32o_fixed_mul(int32_t a, int32_t b);                     /* fixed.rs:161-165 */
int32_t b32o_div_unr(int32_t num, int32_t den);                   /* fixed.rs:178-230 */
void    b32o_project_fixed(const float world[3], const b32_camera* cam, uint32_t w, uint32_t h,
                           int32_t* sx, int32_t* sy, float* depth); /* fixed.rs:424-441 */

/* render.rs helpers */
void b32o_dither_and_quantize(uint8_t r8, uint8_t g8, uint8_t b8, uint32_t x, uint32_t y,
                              uint8_t out5[3]);                   /* render.rs:1173-1182 */
void b32o_blend_rgb555(uint8_t fr, uint8_t fg, uint8_t fb, uint8_t br, uint8_t bg, uint8_t bb,
                       uint32_t mode, uint8_t out8[3]);           /* render.rs:1093-1145 */
uint16_t b32o_texture_sample(const b32_tex_desc* tex, float u, float v); /* types.rs:671-681 */
void b32o_shade_multi_light(const float normal[3], const float world_pos[3],
                            const b32_light* lights, uint32_t n_lights, float ambient,
                            float out_rgb[3]);                    /* render.rs:1013-1071 */
void b32o_acosf(const float* x, float* out, uint32_t n);   /* f32::acos of the reference's wasm build (compiler_builtins libm acosf), render.rs:1047 */

/* Framebuffer::clear, render.rs:36-45 */
void b32o_fb_clear(uint8_t* rgba, float* z, uint32_t w, uint32_t h,
                   uint8_t r, uint8_t g, uint8_t b, uint8_t a);

/* transform + snap loop, render.rs:2321-2360. out_screen/out_cam: nv*3 floats each. */
void b32o_transform(const b32_vertex* v, uint32_t nv, const b32_camera* cam,
                    const b32_settings* s, uint32_t w, uint32_t h,
                    float* out_screen, float* out_cam);

/* render_mesh_15, render.rs:2302-2572 (wireframe phase :2574-2635 included when enabled).
 * draw_order (nullable): Surface.face_idx in draw order, up to cap entries; *n_drawn = count.
 * Returns a B32_* code (B32_ERR_OOB_INDEX / B32_ERR_NAN_DEPTH where the reference panics). */
int b32o_render_mesh_15(uint8_t* fb_rgba, float* fb_z, uint32_t w, uint32_t h,
                        const b32_vertex* vertices, uint32_t nv,
                        const b32_face* faces, uint32_t nf,
                        const b32_tex_desc* textures, uint32_t ntex,
                        const b32_camera* camera, const b32_settings* settings,
                        const b32_fog* fog_or_null, b32_timings* timings,
                        uint32_t* draw_order, uint32_t cap, uint32_t* n_drawn);

/* render_mesh, render.rs:1971-2259 (RGB888 path: rasterize_triangle render.rs:1202-1433, Color ops
 * types.rs:783-934, Texture::sample types.rs:1242-1253).  Same conventions as b32o_render_mesh_15. */
int b32o_render_mesh(uint8_t* fb_rgba, float* fb_z, uint32_t w, uint32_t h,
                     const b32_vertex* vertices, uint32_t nv,
                     const b32_face* faces, uint32_t nf,
                     const b32_tex8_desc* textures, uint32_t ntex,
                     const b32_camera* camera, const b32_settings* settings, b32_timings* timings,
                     uint32_t* draw_order, uint32_t cap, uint32_t* n_drawn);

/* Framebuffer::render_skybox step 1 (sphere pass), render.rs:89-139 + rasterize_skybox_triangle :242-299. */
int b32o_render_skybox_mesh(uint8_t* fb_rgba, uint32_t w, uint32_t h,
                            const b32_sky_vertex* vertices, uint32_t nv, const uint32_t* faces, uint32_t nf,
                            const b32_camera* camera);

/* render_stars from the star direction on + draw_star_diamond (render.rs:175-235). */
int b32o_render_stars(uint8_t* fb_rgba, uint32_t w, uint32_t h, const b32_star* stars, uint32_t n,
                      const b32_camera* camera, float size);
/* render_asset_parts' vertex transform (src/scene.rs:121-160). */
int b32o_place_vertices(const b32_vertex* in, uint32_t nv, float facing, float cos_f, float sin_f,
                        const float world_pos[3], b32_vertex* out);
/* Framebuffer::clear_gradient (render.rs:60-77). */
void b32o_fb_clear_gradient(uint8_t* rgba, float* z, uint32_t w, uint32_t h,
                            const uint8_t top[3], const uint8_t bottom[3], uint8_t a);
/* Framebuffer::draw_line* (render.rs:684-872), one call per list entry, in order. */
int b32o_draw_lines(uint8_t* fb_rgba, float* fb_z, uint32_t w, uint32_t h, const b32_line* lines, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
