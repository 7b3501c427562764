/*
 * b32_raster.h — C ABI of the B200-native BONNIE-32 rasterizer hot path.
 *
 * The reference (EBonura/bonnie-32) has no FFI around its rasterizer: the boundary is the
 * in-process Rust signature
 *
 *     pub fn render_mesh_15(fb: &mut Framebuffer, vertices: &[Vertex], faces: &[Face],
 *                           textures: &[Texture15], camera: &Camera, settings: &RasterSettings,
 *                           fog: Option<(f32, f32, f32, Color)>) -> RasterTimings
 *                                                   (src/rasterizer/render.rs:2302-2310)
 *
 * plus `struct Framebuffer` (src/rasterizer/render.rs:10-45).  This header is what a Rust
 * `extern "C"` shim keeping those signatures binds (see INTEGRATION.md for the shim).
 * Plain pointers and sizes only; no CUDA or torch types appear in any signature.
 *
 * Threading: a context is not thread-safe; use one context per host thread / GPU.
 * Ownership: the caller owns every buffer it passes; the library borrows it for the call only.
 */
#ifndef B32_RASTER_H
#define B32_RASTER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes (the reference panics where these are returned) --------------------- */
#define B32_OK               0
#define B32_ERR_INVALID      1  /* null pointer / zero-sized framebuffer / bad enum            */
#define B32_ERR_OOB_INDEX    2  /* face.v* >= nv: reference panics on slice index              */
#define B32_ERR_NAN_DEPTH    3  /* NaN sort key: reference `partial_cmp().unwrap()` panics,    */
                                /* src/rasterizer/render.rs:2531                                */
#define B32_ERR_UNSUPPORTED  4  /* a size beyond what the device layout holds (framebuffer side  */
                                /* > 65535, > 65535 textures, ...); a wireframe edge longer than */
                                /* 2^24 pixels (frame otherwise drawn)                          */
#define B32_ERR_CUDA         5  /* a CUDA runtime call failed; see b32_last_error()            */
#define B32_ERR_NO_DEVICE    6  /* no CUDA device: there is NO CPU fallback                    */

/* ---- enums (values = declaration order of the Rust enums) ----------------------------- */
/* BlendMode, src/rasterizer/types.rs:1378-1388 */
enum { B32_BLEND_OPAQUE = 0, B32_BLEND_AVERAGE = 1, B32_BLEND_ADD = 2,
       B32_BLEND_SUBTRACT = 3, B32_BLEND_ADD_QUARTER = 4, B32_BLEND_ERASE = 5 };
/* ShadingMode, src/rasterizer/types.rs:1288-1293 */
enum { B32_SHADE_NONE = 0, B32_SHADE_FLAT = 1, B32_SHADE_GOURAUD = 2 };
/* LightType, src/rasterizer/types.rs:1296-1304 */
enum { B32_LIGHT_DIRECTIONAL = 0, B32_LIGHT_POINT = 1, B32_LIGHT_SPOT = 2 };
/* texture storage formats accepted by b32_textures_set */
enum { B32_TEX_RGB555 = 0,   /* Texture15.pixels: u16 sRRRRRGGGGGBBBBB  (types.rs:532-539)     */
       B32_TEX_IDX8   = 1,   /* one u8 palette index per texel + CLUT   (types.rs:438, 390-397) */
       B32_TEX_IDX4   = 2 }; /* two texels per byte, low nibble = even x, + CLUT               */

/* ---- POD records ---------------------------------------------------------------------- */

/* struct Vertex, src/rasterizer/types.rs:947-959 (bone_index is ignored by the rasterizer).
 * `blend` is Color.blend: it takes part in the `vc1 != vc2` test of needs_dither
 * (render.rs:1487-1492) and is therefore carried.  36 bytes. */
typedef struct b32_vertex {
    float   pos[3];
    float   uv[2];
    float   normal[3];
    uint8_t r, g, b, blend;
} b32_vertex;

/* struct Face, src/rasterizer/types.rs:984-1002.  16 bytes.
 * flags: bits 0-15 texture_id (0xFFFF = None), bits 16-18 blend_mode (a B32_BLEND_* value: 6 and 7 are not a
 *        BlendMode, the call draws nothing and returns B32_ERR_INVALID), bit 19 black_transparent,
 *        bits 24-31 editor_alpha. */
typedef struct b32_face {
    uint32_t v0, v1, v2;
    uint32_t flags;
} b32_face;
#define B32_FACE_TEX_NONE 0xFFFFu
#define B32_FACE_FLAGS(tex_id, blend, black_transparent, editor_alpha)                      \
    (((uint32_t)(tex_id) & 0xFFFFu) | (((uint32_t)(blend) & 7u) << 16) |                    \
     (((uint32_t)((black_transparent) ? 1u : 0u)) << 19) | (((uint32_t)(editor_alpha) & 0xFFu) << 24))

/* struct Camera, src/rasterizer/camera.rs:9-18 (basis vectors are computed on the host by
 * Camera::update_basis, camera.rs:76-91, and passed as data). 48 bytes. */
typedef struct b32_camera {
    float position[3];
    float basis_x[3];
    float basis_y[3];
    float basis_z[3];
} b32_camera;

/* struct Light / enum LightType, src/rasterizer/types.rs:1296-1314.
 * direction must already be normalised (Light::directional does it, types.rs:1319). */
typedef struct b32_light {
    uint32_t type;           /* B32_LIGHT_*                                   */
    float    position[3];    /* Point, Spot                                   */
    float    direction[3];   /* Directional, Spot                             */
    float    radius;         /* Point, Spot                                   */
    float    angle;          /* Spot: cone half-angle in radians; `acos` of render.rs:1047   */
                             /* is the libm `acosf` of the reference's shipped wasm build    */
    float    intensity;
    uint8_t  r, g, b;        /* Light.color                                   */
    uint8_t  enabled;
} b32_light;

/* struct RasterSettings, src/rasterizer/types.rs:1392-1428 (low_resolution and stretch_to_fill
 * are UI-only and not carried). */
typedef struct b32_settings {
    uint8_t affine_textures;
    uint8_t use_zbuffer;
    uint8_t shading;             /* B32_SHADE_*                               */
    uint8_t backface_cull;
    uint8_t backface_wireframe;
    uint8_t dithering;
    uint8_t wireframe_overlay;
    uint8_t use_rgb555;          /* informational: callers pick render_mesh vs render_mesh_15 */
    uint8_t use_fixed_point;
    uint8_t xray_mode;
    uint8_t ortho_enabled;       /* ortho_projection.is_some()                */
    uint8_t _pad;
    float   ambient;
    float   ortho_zoom, ortho_center_x, ortho_center_y;
    uint32_t        n_lights;
    const b32_light* lights;     /* n_lights entries, may be NULL when 0      */
} b32_settings;

/* fog: Option<(start, falloff, cull_distance, Color)>, render.rs:2309 */
typedef struct b32_fog {
    float   start, falloff, cull_distance;
    uint8_t r, g, b, blend;
} b32_fog;

/* struct RasterTimings, src/rasterizer/types.rs:1499-1514.  Phase times are device times
 * (CUDA events) of the kernels that replace each reference phase. */
typedef struct b32_timings {
    float    transform_ms, fog_ms, cull_ms, sort_ms, draw_ms, wireframe_ms;
    uint32_t triangles_drawn;
} b32_timings;

/* struct Texture15 (types.rs:532-539) or an indexed texture + CLUT (types.rs:438, 328-397;
 * IndexedAtlas::to_texture15, src/modeler/mesh_editor.rs:669-682). */
typedef struct b32_tex_desc {
    uint32_t        width, height;
    uint32_t        format;      /* B32_TEX_*                                 */
    uint32_t        blend_mode;  /* Texture15.blend_mode                      */
    const void*     pixels;      /* u16[w*h] | u8[w*h] | u8[(w*h+1)/2]        */
    const uint16_t* clut;        /* indexed formats: clut_len Color15 entries */
    uint32_t        clut_len;    /* index >= clut_len samples 0x0000 (Clut::lookup, types.rs:390-397) */
} b32_tex_desc;

/* struct Texture of the RGB888 path (types.rs:1058-1066): `pixels: Vec<Color>`, one Color per texel
 * marshalled as 4 bytes r, g, b, blend (struct Color, types.rs:721-726; blend = B32_BLEND_*, Erase =
 * transparent texel, types.rs:783-785). */
typedef struct b32_tex8_desc {
    uint32_t       width, height;
    uint32_t       blend_mode;   /* Texture.blend_mode: only feeds the unused has_transparency (render.rs:2074-2078) */
    uint32_t       _pad;
    const uint8_t* pixels;       /* 4 * w * h bytes                          */
} b32_tex8_desc;

typedef struct b32_ctx  b32_ctx;   /* one GPU, one stream, one device-resident Framebuffer   */
typedef struct b32_mesh b32_mesh;  /* device-resident vertex/face buffers                    */

/* ---- context -------------------------------------------------------------------------- */
int         b32_ctx_create(int device, b32_ctx** out);
void        b32_ctx_destroy(b32_ctx* ctx);
const char* b32_last_error(const b32_ctx* ctx);
/* The cudaStream_t all work of this context is enqueued on (for external event timing). */
void*       b32_ctx_stream(b32_ctx* ctx);
/* Block until everything enqueued so far has finished; returns the first deferred error. */
int         b32_sync(b32_ctx* ctx);
/* Number of kernels of this library launched on this context so far. */
uint64_t    b32_kernel_launches(const b32_ctx* ctx);

/* ---- Framebuffer (src/rasterizer/render.rs:10-77) -------------------------------------- */
/* Framebuffer::new / resize: pixels zeroed, zbuffer = f32::MAX (render.rs:18-34). */
int b32_fb_resize(b32_ctx* ctx, uint32_t width, uint32_t height);
/* Framebuffer::clear(color): every pixel = (r,g,b,a), zbuffer = f32::MAX (render.rs:36-45). */
int b32_fb_clear(b32_ctx* ctx, uint8_t r, uint8_t g, uint8_t b, uint8_t a);
/* Framebuffer::clear_gradient(top, bottom) (render.rs:60-77): row y gets Color::lerp(top, bottom, y/(h-1))
 * (types.rs:811-820, f32, truncating casts; t = 0 when h == 1), zbuffer = f32::MAX.  a = the alpha byte of
 * Color::to_bytes for `top` (255, or 0 when top.blend == Erase; lerp keeps top's blend tag). */
int b32_fb_clear_gradient(b32_ctx* ctx, uint8_t top_r, uint8_t top_g, uint8_t top_b,
                          uint8_t bottom_r, uint8_t bottom_g, uint8_t bottom_b, uint8_t a);
/* Host access to Framebuffer.pixels / .zbuffer (host overlays, present). z may be NULL. */
int b32_fb_upload(b32_ctx* ctx, const uint8_t* rgba, const float* z);
int b32_fb_download(b32_ctx* ctx, uint8_t* rgba, float* z);
int b32_fb_size(const b32_ctx* ctx, uint32_t* width, uint32_t* height);

/* ---- textures (&[Texture15] argument of render_mesh_15) -------------------------------- */
/* Replaces the context's texture table; face.texture_id indexes it. Cached across calls. */
int b32_textures_set(b32_ctx* ctx, const b32_tex_desc* descs, uint32_t n);

/* ---- the hot path ---------------------------------------------------------------------- */
/* render_mesh_15 (render.rs:2302-2638) on host buffers: copies vertices/faces to the device,
 * renders into the context's framebuffer, waits, fills *timings (may be NULL). */
int b32_render_mesh_15(b32_ctx* ctx,
                       const b32_vertex* vertices, uint32_t nv,
                       const b32_face* faces, uint32_t nf,
                       const b32_camera* camera, const b32_settings* settings,
                       const b32_fog* fog_or_null, b32_timings* timings);

/* Same call with control flags.  B32_RENDER_ASYNC: copy + render are only enqueued on the context's
 * stream (the host buffers must stay valid and unchanged until b32_sync / b32_fb_download returns);
 * no timings; errors surface at b32_sync / b32_fb_download.  Both passes of render_mesh_15 are enqueued
 * (render.rs:2551-2569): the semi-transparent pass decides on the device whether it has anything to do.
 * B32_RENDER_ALL_OPAQUE (optional) is the caller's promise that there is no second pass — no face has
 * blend_mode != Opaque or editor_alpha < 255, no bound texture has blend_mode != Opaque, x-ray mode is
 * off (a marshalling shim knows this for free) — which saves its launch; the device checks the promise
 * and reports B32_ERR_INVALID at the next sync if it was wrong.  The wireframe phase cannot be enqueued.
 * One limit: a screen tile holding more than 2048 semi-transparent surfaces needs scratch memory that
 * only a blocking call sizes; an enqueued call that meets such a tile before any blocking call did
 * reports B32_ERR_UNSUPPORTED at the next sync (pass 2 was not drawn). */
#define B32_RENDER_ASYNC       1u
#define B32_RENDER_ALL_OPAQUE  2u
/* Compact marshalling for the host-buffer path (the bytes that cross PCIe every call; results are identical):
 * B32_VTX_NO_NORMAL   `vertices` points to b32_vertex_nn records (24 bytes: no normal).  Legal only with
 *                     settings.shading == None — the normals are read by nothing else (render.rs:1466-1483);
 *                     otherwise B32_ERR_INVALID.
 * B32_FACES_IMPLICIT  `faces` points to nf uint32_t flags words (the `flags` of b32_face); face i uses vertices
 *                     3i, 3i+1, 3i+2 (an unindexed triangle soup); nv >= 3 * nf, else B32_ERR_OOB_INDEX.
 * B32_FACES_UNIFORM   an unindexed soup (as above) whose faces all carry the same flags word — one texture, one blend
 *                     mode, the usual case of a room or asset part: `faces` points to that ONE uint32_t; it travels with
 *                     the kernel parameters, no face buffer is copied at all.
 * A marshalling shim picks them for free while it converts `&[Vertex]` / `&[Face]`: 36 + 16/3 -> 24 + 4/3 (or 24) bytes per vertex. */
#define B32_VTX_NO_NORMAL      4u
#define B32_FACES_IMPLICIT     8u
#define B32_FACES_UNIFORM      16u
typedef struct b32_vertex_nn {
    float   pos[3];
    float   uv[2];
    uint8_t r, g, b, blend;
} b32_vertex_nn;
int b32_render_mesh_15_ex(b32_ctx* ctx,
                          const void* vertices, uint32_t nv,      /* b32_vertex[nv], or b32_vertex_nn[nv] */
                          const void* faces, uint32_t nf,         /* b32_face[nf], uint32_t flags[nf], or one uint32_t */
                          const b32_camera* camera, const b32_settings* settings,
                          const b32_fog* fog_or_null, uint32_t flags, b32_timings* timings);
/* Enqueue the framebuffer read-back (pinned destination recommended); complete after b32_sync. */
int b32_fb_download_async(b32_ctx* ctx, uint8_t* rgba, float* z);

/* Device-resident geometry: upload once, render many times (static level geometry; this is
 * also how `value` is measured with inputs already in HBM). */
int  b32_mesh_upload(b32_ctx* ctx, const b32_vertex* vertices, uint32_t nv,
                     const b32_face* faces, uint32_t nf, b32_mesh** out);
void b32_mesh_free(b32_ctx* ctx, b32_mesh* mesh);
int  b32_render_mesh_15_resident(b32_ctx* ctx, const b32_mesh* mesh,
                                 const b32_camera* camera, const b32_settings* settings,
                                 const b32_fog* fog_or_null, b32_timings* timings);
/* Same, but only enqueues (no wait, no timings); errors surface at b32_sync/b32_fb_download.  Both passes
 * are enqueued; only a call with the wireframe phase on is rendered synchronously instead. */
int  b32_render_mesh_15_enqueue(b32_ctx* ctx, const b32_mesh* mesh,
                                const b32_camera* camera, const b32_settings* settings,
                                const b32_fog* fog_or_null);

/* Framebuffer::clear(color) + render_mesh_15 on resident geometry as ONE enqueued frame (clear_rgba = 4 bytes
 * r, g, b, a; NULL = no clear).  From the third frame of one mesh / framebuffer size on, the frame is replayed as
 * a CUDA graph whose kernel nodes are re-parameterised in place (camera, settings, lights, fog change freely):
 * one driver call per frame instead of one per kernel.  b32_render_mesh_15_enqueue takes the same path. */
int  b32_frame_15_enqueue(b32_ctx* ctx, const uint8_t* clear_rgba, const b32_mesh* mesh,
                          const b32_camera* camera, const b32_settings* settings, const b32_fog* fog_or_null);
/* Number of frames launched as graphs on this context so far. */
uint64_t b32_graph_launches(const b32_ctx* ctx);
/* RasterTimings for frames that were only enqueued (the debug overlay of src/game/renderer.rs:735-980 wants them every
 * frame, and an enqueue has nothing to return yet).  Opt-in per context, because the frame's kernels then publish
 * their counters and %globaltimer stamps in host-mapped memory (one atomic per CTA).  b32_frame_timings never blocks: it
 * fills `out` from the most recent enqueued frame that has finished (all zeros while none has).  cull_ms = the fused
 * transform + cull + setup kernel, draw_ms = pass 1 + pass 2, triangles_drawn exact; the other fields stay 0. */
int b32_ctx_frame_timings(b32_ctx* ctx, int enable);
int b32_frame_timings(b32_ctx* ctx, b32_timings* out);

/* Placed asset parts: render_asset_parts (src/scene.rs:109-169) draws one mesh part per render_mesh* call after
 * rotating its vertices about Y by the object's `facing` and translating them to `world_pos` on the host, every frame.
 * Here the part is resident (b32_mesh_upload, once) and the per-object transform runs on the device:
 *   pos    = (x * cos_f - z * sin_f + wx,  y + wy,  x * sin_f + z * cos_f + wz)
 *   normal = (nx * cos_f - nz * sin_f,     ny,      nx * sin_f + nz * cos_f)        (scene.rs:141-160)
 * cos_f / sin_f are the caller's `facing.cos()` / `facing.sin()` (libm stays on the host).  As in the reference the
 * transform is skipped when |facing| and every |world_pos| component are <= 0.0001.  flags: 0 = blocking with timings,
 * B32_RENDER_ASYNC = enqueue only (like b32_render_mesh_15_enqueue).  rgb888 != 0 selects render_mesh (fog ignored). */
typedef struct b32_placement {
    float facing, cos_f, sin_f;
    float world_pos[3];
} b32_placement;
int  b32_render_mesh_placed(b32_ctx* ctx, const b32_mesh* mesh, const b32_placement* placement,
                            const b32_camera* camera, const b32_settings* settings, const b32_fog* fog_or_null,
                            int rgb888, uint32_t flags, b32_timings* timings);

/* ---- the RGB888 sibling (RasterSettings.use_rgb555 == false) ----------------------------- */
/* `textures: &[Texture]` of render_mesh: a second texture table, independent of b32_textures_set's. */
int b32_textures_set_rgb888(b32_ctx* ctx, const b32_tex8_desc* descs, uint32_t n);
/* render_mesh (render.rs:1971-2259) -> rasterize_triangle (render.rs:1202-1433): no fog, ONE surface list
 * (sorted back to front only when !use_zbuffer), 8-bit colour pipeline, per-texel blend tags, 8-bit blends
 * (Color::blend_with, types.rs:886-930), dither re-expanded with `<< 3` (render.rs:1186-1197). */
int b32_render_mesh(b32_ctx* ctx,
                    const b32_vertex* vertices, uint32_t nv,
                    const b32_face* faces, uint32_t nf,
                    const b32_camera* camera, const b32_settings* settings, b32_timings* timings);
int b32_render_mesh_resident(b32_ctx* ctx, const b32_mesh* mesh,
                             const b32_camera* camera, const b32_settings* settings, b32_timings* timings);

/* ---- skybox sphere pass (Framebuffer::render_skybox, render.rs:81-139) -------------------------------- */
/* One vertex of Skybox::generate_mesh (src/world/geometry.rs:529-, struct SkyboxVertex :1027-1030): world position
 * + colour.  The mesh (sphere + mountains, libm sin/cos/powf) is generated on the host as in the reference. */
typedef struct b32_sky_vertex {
    float   pos[3];
    uint8_t r, g, b, _pad;
} b32_sky_vertex;
/* Step 1 of render_skybox: float transform + `project` of every vertex (behind-camera vertices drop their faces),
 * inward-facing triangles only (signed area < 0), rasterize_skybox_triangle (render.rs:242-299): pixel centres,
 * Gouraud vertex colours, no depth test or write, faces drawn in order (later ones overwrite).  faces = 3*nf
 * vertex indices.  Step 2 is b32_render_stars.  Skies whose worst-case tile bins fit 256 MB (any realistic one) are
 * enqueued without a wait; the lists are consumed before the call returns. */
int b32_render_skybox_mesh(b32_ctx* ctx, const b32_sky_vertex* vertices, uint32_t nv,
                           const uint32_t* faces, uint32_t nf, const b32_camera* camera);

/* Step 2 of render_skybox: render_stars + draw_star_diamond (render.rs:149-235).  The host keeps what needs libm and
 * the LCG: for star i (in the reference's loop order, visible or not) dir = (sin(phi)cos(theta), cos(phi),
 * sin(phi)sin(theta)) and r, g, b = (stars.color.c as f32 * brightness) as u8.  The device does `dir * 10000.0`,
 * perspective_transform, the `cam_space.z > 0.1` test, `project`, and the diamond of 1 / 5 / 9 set_pixel calls for
 * `size.max(1.0) as i32` = 1 / 2 / >= 3, later stars over earlier ones.  (The twinkle phase is only drawn from the
 * LCG for visible stars, render.rs:186-190, so the host evaluates the same visibility test while it builds the list;
 * it may pass every star or only the visible ones.) */
typedef struct b32_star {
    float   dir[3];
    uint8_t r, g, b, _pad;
} b32_star;
/* Enqueued without a wait; the list is consumed before the call returns. */
int b32_render_stars(b32_ctx* ctx, const b32_star* stars, uint32_t n, const b32_camera* camera, float size);

/* ---- overlay lines (Framebuffer::draw_line*, render.rs:684-872) --------------------------------------- */
/* The post-passes the editor and the game draw over a rendered frame (grids, gizmos, wireframes, collision
 * shapes).  On the host they force a framebuffer download between the render and the present; here a whole
 * list is drawn on the device, with the result of calling the Framebuffer methods one by one in list order. */
enum {
    B32_LINE_2D          = 0,   /* draw_line / draw_line_blended(mode) (:714-751): set_pixel or set_pixel_blended */
    B32_LINE_2D_ALPHA    = 1,   /* draw_line_alpha (:684-711): set_pixel_alpha (:646-667) */
    B32_LINE_3D          = 2,   /* draw_line_3d (:756-758): depth test z < zbuffer, set_pixel */
    B32_LINE_3D_OVERLAY  = 3,   /* draw_line_3d_overlay (:763-765): z <= zbuffer, set_pixel */
    B32_LINE_3D_ALPHA    = 4,   /* draw_line_3d_alpha (:822-872): z * 0.995 <= zbuffer, set_pixel_alpha */
    /* the filled primitives of the same family ride in the same list (and the same ordering): */
    B32_LINE_CIRCLE       = 5,  /* draw_circle (:631-644): x0, y0 = centre, x1 = radius (|radius| <= 32767); set_pixel */
    B32_LINE_CIRCLE_ALPHA = 6,  /* draw_circle_alpha (:670-682): the same with set_pixel_alpha */
    B32_LINE_FILLED_RECT  = 7,  /* draw_filled_rect (:954-972): corners (x0, y0), (x1, y1) inclusive; set_pixel.
                                   draw_rect (:941-951) is four B32_LINE_2D entries: top, right, bottom, left */
    B32_LINE_THICK        = 8   /* draw_thick_line (:875-938): z0 = thickness as f32 (an integer); thickness <= 1 is draw_line */
};
typedef struct b32_line {
    int32_t x0, y0, x1, y1;     /* end points, |coordinate| <= B32_LINE_MAX_COORD */
    float   z0, z1;             /* end-point depths of the 3D kinds (never written to the z-buffer) */
    uint8_t r, g, b, blend;     /* struct Color (types.rs:721-726); blend only decides to_bytes' alpha (Erase: 0) */
    uint8_t kind;               /* B32_LINE_* */
    uint8_t mode;               /* B32_LINE_2D: the BlendMode argument of draw_line_blended (B32_BLEND_OPAQUE = draw_line) */
    uint8_t alpha;              /* the *_ALPHA kinds */
    uint8_t _pad;
} b32_line;
#define B32_LINE_MAX_COORD (1 << 20)
/* The list is consumed before the call returns.  Lists of overwriting lines only (draw_line, draw_line_3d,
 * draw_line_3d_overlay, Erase) are enqueued without a wait; lists with blended lines wait once per batch of rounds.
 * B32_ERR_INVALID for an unknown kind/mode, B32_ERR_UNSUPPORTED for a coordinate beyond
 * B32_LINE_MAX_COORD (the reference walks every step of a line, on screen or not). */
int b32_draw_lines(b32_ctx* ctx, const b32_line* lines, uint32_t n);

/* ---- pinned host memory for callers that want zero-copy DMA of their Vec buffers -------- */
void* b32_host_alloc(size_t bytes);
void  b32_host_free(void* p);

/* ---- per-stage device outputs, for parity tests of the individual kernels --------------- */
/* Transform + snap only (render.rs:2321-2360): out_screen[nv*3] = projected (x,y,z),
 * out_cam[nv*3] = cam_space_positions. */
int b32_debug_transform(b32_ctx* ctx, const b32_vertex* vertices, uint32_t nv,
                        const b32_camera* camera, const b32_settings* settings,
                        float* out_screen, float* out_cam);
/* Draw order of the last render call: out_face_idx[i] = Surface.face_idx of the i-th surface
 * drawn (opaque pass then transparent pass). Returns count via *n (<= cap written). */
int b32_debug_draw_order(b32_ctx* ctx, uint32_t* out_face_idx, uint32_t cap, uint32_t* n);

/* Device time (ms, CUDA events on the context stream) of each kernel group of the last synchronous
 * render call: [0] k_setup (transform + cull + setup + tile masks) [1] k_fill_opaque (pass 1),
 * [2] unused (0) [3] k_fill_ordered (pass 2 / x-ray: per-tile sort + replay).
 * Returns the number of values written (<= cap). */
int b32_debug_kernel_times(b32_ctx* ctx, float* out_ms, uint32_t cap);
/* Measurement aid for enqueued frames (bench.py's per-kernel times in the regime of its timed loop): with n > 0 the
 * context's next enqueued frames are launched plainly (no CUDA graph) with events in front of k_setup, in front of the
 * fill kernel(s) and behind them, n frames deep; n = 0 switches it off.  b32_debug_timing_read (syncs) returns the
 * number of frames read; setup_ms[i] / fill_ms[i] are their device times. */
int b32_debug_timing_ring(b32_ctx* ctx, uint32_t n);
/* 1 when the next fixed-point call on this context will run the shared-edge-prefix fill because one of its last 8 calls
 * (as far as the device has got) drew a large surface with stepped (non-integer-exact) edge values; 0 otherwise; < 0 on error. */
int b32_debug_prefix_hint(b32_ctx* ctx);
int b32_debug_timing_read(b32_ctx* ctx, float* setup_ms, float* fill_ms, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* B32_RASTER_H */
