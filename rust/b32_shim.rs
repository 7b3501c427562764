//! b32_shim.rs — drop-in replacement body for `render_mesh_15` that forwards to the CUDA library.
//!
//! NOT compiled in this repository's environment (no rustc); it is the reference-side binding a
//! BONNIE-32 maintainer adds.  Put it at `src/rasterizer/b32_shim.rs`, add `mod b32_shim;` to
//! `src/rasterizer/mod.rs`, and re-export `b32_shim::render_mesh_15` instead of
//! `render::render_mesh_15` (mod.rs:63).  Link with `-lb32raster` (build.rs:
//! `println!("cargo:rustc-link-lib=dylib=b32raster")`).
//!
//! Signature and semantics are those of `src/rasterizer/render.rs:2302-2310`.  Where the reference
//! panics (bad vertex index, NaN sort key) this shim panics with the library's message.
#![allow(non_camel_case_types)]
use super::camera::Camera;
use super::render::Framebuffer;
use super::types::{BlendMode, Color, Face, LightType, RasterSettings, RasterTimings, ShadingMode, Texture15, Vertex};
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] #[derive(Clone, Copy)]
pub struct b32_vertex { pos: [f32; 3], uv: [f32; 2], normal: [f32; 3], r: u8, g: u8, b: u8, blend: u8 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct b32_face { v0: u32, v1: u32, v2: u32, flags: u32 }
#[repr(C)] pub struct b32_camera { position: [f32; 3], basis_x: [f32; 3], basis_y: [f32; 3], basis_z: [f32; 3] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct b32_light { kind: u32, position: [f32; 3], direction: [f32; 3], radius: f32, angle: f32, intensity: f32,
                       r: u8, g: u8, b: u8, enabled: u8 }
#[repr(C)] pub struct b32_settings {
    affine_textures: u8, use_zbuffer: u8, shading: u8, backface_cull: u8, backface_wireframe: u8, dithering: u8,
    wireframe_overlay: u8, use_rgb555: u8, use_fixed_point: u8, xray_mode: u8, ortho_enabled: u8, _pad: u8,
    ambient: f32, ortho_zoom: f32, ortho_center_x: f32, ortho_center_y: f32, n_lights: u32, lights: *const b32_light }
#[repr(C)] pub struct b32_fog { start: f32, falloff: f32, cull_distance: f32, r: u8, g: u8, b: u8, blend: u8 }
#[repr(C)] #[derive(Default)]
pub struct b32_timings { transform_ms: f32, fog_ms: f32, cull_ms: f32, sort_ms: f32, draw_ms: f32, wireframe_ms: f32,
                         triangles_drawn: u32 }
#[repr(C)] pub struct b32_tex_desc { width: u32, height: u32, format: u32, blend_mode: u32, pixels: *const c_void,
                                     clut: *const u16, clut_len: u32 }
pub enum b32_ctx {}

extern "C" {
    fn b32_ctx_create(device: c_int, out: *mut *mut b32_ctx) -> c_int;
    fn b32_last_error(ctx: *const b32_ctx) -> *const c_char;
    fn b32_fb_resize(ctx: *mut b32_ctx, w: u32, h: u32) -> c_int;
    fn b32_fb_upload(ctx: *mut b32_ctx, rgba: *const u8, z: *const f32) -> c_int;
    fn b32_fb_download(ctx: *mut b32_ctx, rgba: *mut u8, z: *mut f32) -> c_int;
    fn b32_textures_set(ctx: *mut b32_ctx, descs: *const b32_tex_desc, n: u32) -> c_int;
    fn b32_render_mesh_15_ex(ctx: *mut b32_ctx, v: *const c_void, nv: u32, f: *const c_void, nf: u32, cam: *const b32_camera,
                             s: *const b32_settings, fog: *const b32_fog, flags: u32, tm: *mut b32_timings) -> c_int;
    fn b32_render_mesh_15(ctx: *mut b32_ctx, v: *const b32_vertex, nv: u32, f: *const b32_face, nf: u32,
                          cam: *const b32_camera, s: *const b32_settings, fog: *const b32_fog,
                          out: *mut b32_timings) -> c_int;
}

fn blend_u8(b: BlendMode) -> u8 { b as u8 }   // declaration order = B32_BLEND_* (types.rs:1378-1388)

thread_local! {
    // one context per (main) thread: the reference renders on the macroquad main thread only
    static CTX: *mut b32_ctx = unsafe {
        let mut c = std::ptr::null_mut();
        assert_eq!(b32_ctx_create(0, &mut c), 0, "b32_ctx_create failed: no CUDA device (there is no CPU fallback)");
        c
    };
    static TEX_GEN: std::cell::Cell<(usize, usize)> = std::cell::Cell::new((0, 0));
}

fn check(ctx: *mut b32_ctx, rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(b32_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("b32 rasterizer error {}: {}", rc, msg);   // the reference panics in the same situations
    }
}

#[repr(C)] #[derive(Clone, Copy)]
pub struct b32_vertex_nn { pos: [f32; 3], uv: [f32; 2], r: u8, g: u8, b: u8, blend: u8 }
const B32_VTX_NO_NORMAL: u32 = 4;
const B32_FACES_IMPLICIT: u32 = 8;
const B32_FACES_UNIFORM: u32 = 16;

fn face_flags(f: &Face) -> u32 {
    let tex = match f.texture_id { Some(id) if id < 0xFFFF => id as u32, _ => 0xFFFF };
    tex | ((blend_u8(f.blend_mode) as u32) << 16) | ((f.black_transparent as u32) << 19) | ((f.editor_alpha as u32) << 24)
}

/// A pinned (page-locked) staging buffer that only grows: the marshalling below writes straight into it, so the
/// upload is one DMA from where the bytes already are (b32_host_alloc, include/b32_raster.h) instead of a second copy
/// through the library's staging ring.
struct Pinned { ptr: *mut u8, cap: usize }
impl Pinned {
    const fn new() -> Self { Pinned { ptr: std::ptr::null_mut(), cap: 0 } }
    fn reserve(&mut self, bytes: usize) -> *mut u8 {
        extern "C" { fn b32_host_alloc(bytes: usize) -> *mut c_void; fn b32_host_free(p: *mut c_void); }
        if bytes > self.cap {
            unsafe {
                if !self.ptr.is_null() { b32_host_free(self.ptr as *mut c_void); }
                self.cap = bytes.next_power_of_two().max(1 << 16);
                self.ptr = b32_host_alloc(self.cap) as *mut u8;
                assert!(!self.ptr.is_null(), "b32_host_alloc failed");
            }
        }
        self.ptr
    }
}
thread_local! {
    static STAGE_V: std::cell::RefCell<Pinned> = std::cell::RefCell::new(Pinned::new());
    static STAGE_F: std::cell::RefCell<Pinned> = std::cell::RefCell::new(Pinned::new());
}

/// The bytes that cross PCIe, in the most compact layout the call allows (include/b32_raster.h, B32_VTX_NO_NORMAL /
/// B32_FACES_IMPLICIT / B32_FACES_UNIFORM): normals are only read when the settings shade (render.rs:1466-1483), an
/// unindexed triangle soup (face i = vertices 3i, 3i+1, 3i+2) needs no index buffer, and when all its faces carry the
/// same flags word (one texture, one blend mode: the usual room or asset part) that one word is all that is sent.  All
/// of it is decided while the slices are converted anyway, and written once, into pinned memory; the rendered bytes are
/// identical.  Returns (vertex bytes, face bytes, flags); the pointers stay valid until the next call on this thread.
fn marshal_compact(vertices: &[Vertex], faces: &[Face], settings: &RasterSettings) -> (*const c_void, *const c_void, u32) {
    let mut flags = 0u32;
    let vp = STAGE_V.with(|st| {
        let mut st = st.borrow_mut();
        if settings.shading == ShadingMode::None {
            flags |= B32_VTX_NO_NORMAL;
            let out = st.reserve(vertices.len() * 24) as *mut b32_vertex_nn;
            for (i, v) in vertices.iter().enumerate() {
                unsafe { out.add(i).write(b32_vertex_nn { pos: [v.pos.x, v.pos.y, v.pos.z], uv: [v.uv.x, v.uv.y],
                                                            r: v.color.r, g: v.color.g, b: v.color.b, blend: blend_u8(v.color.blend) }); }
            }
            out as *const c_void
        } else {
            let out = st.reserve(vertices.len() * 36) as *mut b32_vertex;
            for (i, v) in vertices.iter().enumerate() { unsafe { out.add(i).write(vertex_record(v)); } }
            out as *const c_void
        }
    });
    let soup = vertices.len() >= 3 * faces.len() && faces.iter().enumerate().all(|(i, f)| f.v0 == 3 * i && f.v1 == 3 * i + 1 && f.v2 == 3 * i + 2);
    let fp = STAGE_F.with(|st| {
        let mut st = st.borrow_mut();
        if soup {
            let out = st.reserve(faces.len().max(1) * 4) as *mut u32;
            let mut uniform = !faces.is_empty();
            for (i, f) in faces.iter().enumerate() {
                let w = face_flags(f);
                unsafe { out.add(i).write(w); uniform = uniform && w == *out; }
            }
            flags |= if uniform { B32_FACES_UNIFORM } else { B32_FACES_IMPLICIT };
            out as *const c_void
        } else {
            let out = st.reserve(faces.len().max(1) * 16) as *mut b32_face;
            for (i, f) in faces.iter().enumerate() { unsafe { out.add(i).write(face_record(f)); } }
            out as *const c_void
        }
    });
    (vp, fp, flags)
}

fn vertex_record(v: &Vertex) -> b32_vertex {
    b32_vertex { pos: [v.pos.x, v.pos.y, v.pos.z], uv: [v.uv.x, v.uv.y], normal: [v.normal.x, v.normal.y, v.normal.z],
                 r: v.color.r, g: v.color.g, b: v.color.b, blend: blend_u8(v.color.blend) }
}
fn face_record(f: &Face) -> b32_face {
    b32_face { v0: f.v0.min(u32::MAX as usize) as u32, v1: f.v1.min(u32::MAX as usize) as u32,
               v2: f.v2.min(u32::MAX as usize) as u32, flags: face_flags(f) }
}

// --- marshal &[Vertex] / &[Face] into the 36 B / 16 B POD records (include/b32_raster.h) ---
fn marshal_geometry(vertices: &[Vertex], faces: &[Face]) -> (Vec<b32_vertex>, Vec<b32_face>) {
    (vertices.iter().map(vertex_record).collect(), faces.iter().map(face_record).collect())
}

/// The returned Vec owns the lights `b32_settings.lights` points at: keep it alive across the call.
fn marshal_settings(settings: &RasterSettings) -> (b32_settings, Vec<b32_light>) {
    let lights: Vec<b32_light> = settings.lights.iter().map(|l| {
        let (kind, position, direction, radius, angle) = match l.light_type {
            LightType::Directional { direction: d } => (0, [0.0; 3], [d.x, d.y, d.z], 0.0, 0.0),
            LightType::Point { position: p, radius } => (1, [p.x, p.y, p.z], [0.0; 3], radius, 0.0),
            LightType::Spot { position: p, direction: d, angle, radius } => (2, [p.x, p.y, p.z], [d.x, d.y, d.z], radius, angle),
        };
        b32_light { kind, position, direction, radius, angle, intensity: l.intensity,
                    r: l.color.r, g: l.color.g, b: l.color.b, enabled: l.enabled as u8 } }).collect();
    let (oe, oz, ox, oy) = match &settings.ortho_projection { Some(o) => (1, o.zoom, o.center_x, o.center_y), None => (0, 0.0, 0.0, 0.0) };
    let s = b32_settings {
        affine_textures: settings.affine_textures as u8, use_zbuffer: settings.use_zbuffer as u8,
        shading: match settings.shading { ShadingMode::None => 0, ShadingMode::Flat => 1, ShadingMode::Gouraud => 2 },
        backface_cull: settings.backface_cull as u8, backface_wireframe: settings.backface_wireframe as u8,
        dithering: settings.dithering as u8, wireframe_overlay: settings.wireframe_overlay as u8,
        use_rgb555: settings.use_rgb555 as u8, use_fixed_point: settings.use_fixed_point as u8,
        xray_mode: settings.xray_mode as u8, ortho_enabled: oe, _pad: 0, ambient: settings.ambient,
        ortho_zoom: oz, ortho_center_x: ox, ortho_center_y: oy, n_lights: lights.len() as u32, lights: lights.as_ptr() };
    (s, lights)
}

fn marshal_camera(camera: &Camera) -> b32_camera {
    b32_camera { position: [camera.position.x, camera.position.y, camera.position.z],
        basis_x: [camera.basis_x.x, camera.basis_x.y, camera.basis_x.z],
        basis_y: [camera.basis_y.x, camera.basis_y.y, camera.basis_y.z],
        basis_z: [camera.basis_z.x, camera.basis_z.y, camera.basis_z.z] }
}

/// `RasterTimings` of the most recent finished frame that was only ENQUEUED (`b32_frame_15_enqueue`, the resident /
/// placed calls with `B32_RENDER_ASYNC`): what the game's debug overlay shows (src/game/renderer.rs:735-980).  Never
/// blocks; all zeros until a frame has finished.  Call `enable_frame_timings(true)` once first.
pub fn enable_frame_timings(on: bool) {
    extern "C" { fn b32_ctx_frame_timings(ctx: *mut b32_ctx, enable: c_int) -> c_int; }
    CTX.with(|&ctx| unsafe { check(ctx, b32_ctx_frame_timings(ctx, on as c_int)) })
}
pub fn frame_timings() -> RasterTimings {
    extern "C" { fn b32_frame_timings(ctx: *mut b32_ctx, out: *mut b32_timings) -> c_int; }
    CTX.with(|&ctx| unsafe {
        let mut tm = b32_timings::default();
        check(ctx, b32_frame_timings(ctx, &mut tm));
        timings(&tm)
    })
}

fn timings(tm: &b32_timings) -> RasterTimings {
    RasterTimings { transform_ms: tm.transform_ms, fog_ms: tm.fog_ms, cull_ms: tm.cull_ms, sort_ms: tm.sort_ms,
                    draw_ms: tm.draw_ms, wireframe_ms: tm.wireframe_ms, triangles_drawn: tm.triangles_drawn }
}

pub fn render_mesh_15(fb: &mut Framebuffer, vertices: &[Vertex], faces: &[Face], textures: &[Texture15],
                      camera: &Camera, settings: &RasterSettings, fog: Option<(f32, f32, f32, Color)>) -> RasterTimings {
    CTX.with(|&ctx| unsafe {
        let (v, f, layout) = marshal_compact(vertices, faces, settings);
        // --- textures: re-upload only when the slice changed (cf. textures_15_cache_generation) ---
        let key = (textures.as_ptr() as usize, textures.len());
        if TEX_GEN.with(|g| g.replace(key)) != key {
            let d: Vec<b32_tex_desc> = textures.iter().map(|t| b32_tex_desc {
                width: t.width as u32, height: t.height as u32, format: 0, blend_mode: blend_u8(t.blend_mode) as u32,
                pixels: t.pixels.as_ptr() as *const c_void, clut: std::ptr::null(), clut_len: 0 }).collect();
            check(ctx, b32_textures_set(ctx, d.as_ptr(), d.len() as u32));
        }
        let (s, _lights) = marshal_settings(settings);
        let cam = marshal_camera(camera);
        let fogc = fog.map(|(start, falloff, cull_distance, c)| b32_fog { start, falloff, cull_distance, r: c.r, g: c.g, b: c.b, blend: blend_u8(c.blend) });
        // --- framebuffer: the host owns fb.pixels / fb.zbuffer between calls (clear, skybox, overlays) ---
        check(ctx, b32_fb_resize(ctx, fb.width as u32, fb.height as u32));
        check(ctx, b32_fb_upload(ctx, fb.pixels.as_ptr(), fb.zbuffer.as_ptr()));
        let mut tm = b32_timings::default();
        check(ctx, b32_render_mesh_15_ex(ctx, v, vertices.len() as u32, f, faces.len() as u32,
                                         &cam, &s, fogc.as_ref().map_or(std::ptr::null(), |f| f as *const _), layout, &mut tm));
        check(ctx, b32_fb_download(ctx, fb.pixels.as_mut_ptr(), fb.zbuffer.as_mut_ptr()));
        timings(&tm)
    })
}

// ---------------------------------------------------------------------------------------------------
// RGB888 sibling: `render_mesh` (src/rasterizer/render.rs:1971-1978), picked by callers when
// `settings.use_rgb555` is false.  Same marshalling; `Texture.pixels: Vec<Color>` becomes 4 bytes per texel.
// ---------------------------------------------------------------------------------------------------
#[repr(C)] pub struct b32_tex8_desc { width: u32, height: u32, blend_mode: u32, _pad: u32, pixels: *const u8 }
extern "C" {
    fn b32_textures_set_rgb888(ctx: *mut b32_ctx, descs: *const b32_tex8_desc, n: u32) -> c_int;
    fn b32_render_mesh(ctx: *mut b32_ctx, v: *const b32_vertex, nv: u32, f: *const b32_face, nf: u32,
                       cam: *const b32_camera, s: *const b32_settings, out: *mut b32_timings) -> c_int;
}
thread_local! { static TEX8_GEN: std::cell::Cell<(usize, usize)> = std::cell::Cell::new((0, 0)); }

pub fn render_mesh(fb: &mut Framebuffer, vertices: &[Vertex], faces: &[Face], textures: &[super::types::Texture],
                   camera: &Camera, settings: &RasterSettings) -> RasterTimings {
    CTX.with(|&ctx| unsafe {
        let (v, f) = marshal_geometry(vertices, faces);
        let key = (textures.as_ptr() as usize, textures.len());
        if TEX8_GEN.with(|g| g.replace(key)) != key {
            let texels: Vec<Vec<u8>> = textures.iter().map(|t| t.pixels.iter()
                .flat_map(|c| [c.r, c.g, c.b, blend_u8(c.blend)]).collect()).collect();
            let d: Vec<b32_tex8_desc> = textures.iter().zip(&texels).map(|(t, px)| b32_tex8_desc {
                width: t.width as u32, height: t.height as u32, blend_mode: blend_u8(t.blend_mode) as u32, _pad: 0,
                pixels: px.as_ptr() }).collect();
            check(ctx, b32_textures_set_rgb888(ctx, d.as_ptr(), d.len() as u32));
        }
        let (s, _lights) = marshal_settings(settings);
        let cam = marshal_camera(camera);
        check(ctx, b32_fb_resize(ctx, fb.width as u32, fb.height as u32));
        check(ctx, b32_fb_upload(ctx, fb.pixels.as_ptr(), fb.zbuffer.as_ptr()));
        let mut tm = b32_timings::default();
        check(ctx, b32_render_mesh(ctx, v.as_ptr(), v.len() as u32, f.as_ptr(), f.len() as u32, &cam, &s, &mut tm));
        check(ctx, b32_fb_download(ctx, fb.pixels.as_mut_ptr(), fb.zbuffer.as_mut_ptr()));
        timings(&tm)
    })
}

// ---------------------------------------------------------------------------------------------------
// Overlay lines: Framebuffer::draw_line* (src/rasterizer/render.rs:684-872) and clear_gradient (:60-77).
// A caller that keeps the framebuffer on the device between the render and the present collects its overlay
// lines in a LineList (same method names and arguments as the Framebuffer methods) and flushes it once: the
// device draws the list with the result of the calls made one after the other.
// ---------------------------------------------------------------------------------------------------
#[repr(C)] #[derive(Clone, Copy)]
pub struct b32_line { x0: i32, y0: i32, x1: i32, y1: i32, z0: f32, z1: f32, r: u8, g: u8, b: u8, blend: u8,
                      kind: u8, mode: u8, alpha: u8, _pad: u8 }
extern "C" {
    fn b32_draw_lines(ctx: *mut b32_ctx, lines: *const b32_line, n: u32) -> c_int;
    fn b32_fb_clear_gradient(ctx: *mut b32_ctx, tr: u8, tg: u8, tb: u8, br: u8, bg: u8, bb: u8, a: u8) -> c_int;
}

#[derive(Default)]
pub struct LineList(Vec<b32_line>);

impl LineList {
    fn push(&mut self, kind: u8, p: (i32, i32, i32, i32), z: (f32, f32), c: Color, mode: BlendMode, alpha: u8) {
        self.0.push(b32_line { x0: p.0, y0: p.1, x1: p.2, y1: p.3, z0: z.0, z1: z.1, r: c.r, g: c.g, b: c.b,
                               blend: blend_u8(c.blend), kind, mode: blend_u8(mode), alpha, _pad: 0 });
    }
    pub fn draw_line(&mut self, x0: i32, y0: i32, x1: i32, y1: i32, color: Color) {
        self.push(0, (x0, y0, x1, y1), (0.0, 0.0), color, BlendMode::Opaque, 255);
    }
    pub fn draw_line_blended(&mut self, x0: i32, y0: i32, x1: i32, y1: i32, color: Color, mode: BlendMode) {
        self.push(0, (x0, y0, x1, y1), (0.0, 0.0), color, mode, 255);
    }
    pub fn draw_line_alpha(&mut self, x0: i32, y0: i32, x1: i32, y1: i32, color: Color, alpha: u8) {
        self.push(1, (x0, y0, x1, y1), (0.0, 0.0), color, BlendMode::Opaque, alpha);
    }
    pub fn draw_line_3d(&mut self, x0: i32, y0: i32, z0: f32, x1: i32, y1: i32, z1: f32, color: Color) {
        self.push(2, (x0, y0, x1, y1), (z0, z1), color, BlendMode::Opaque, 255);
    }
    pub fn draw_line_3d_overlay(&mut self, x0: i32, y0: i32, z0: f32, x1: i32, y1: i32, z1: f32, color: Color) {
        self.push(3, (x0, y0, x1, y1), (z0, z1), color, BlendMode::Opaque, 255);
    }
    pub fn draw_line_3d_alpha(&mut self, x0: i32, y0: i32, z0: f32, x1: i32, y1: i32, z1: f32, color: Color, alpha: u8) {
        self.push(4, (x0, y0, x1, y1), (z0, z1), color, BlendMode::Opaque, alpha);
    }
    pub fn draw_circle(&mut self, cx: i32, cy: i32, radius: i32, color: Color) {
        self.push(5, (cx, cy, radius, 0), (0.0, 0.0), color, BlendMode::Opaque, 255);
    }
    pub fn draw_circle_alpha(&mut self, cx: i32, cy: i32, radius: i32, color: Color, alpha: u8) {
        self.push(6, (cx, cy, radius, 0), (0.0, 0.0), color, BlendMode::Opaque, alpha);
    }
    pub fn draw_filled_rect(&mut self, x0: i32, y0: i32, x1: i32, y1: i32, color: Color) {
        self.push(7, (x0, y0, x1, y1), (0.0, 0.0), color, BlendMode::Opaque, 255);
    }
    pub fn draw_thick_line(&mut self, x0: i32, y0: i32, x1: i32, y1: i32, thickness: i32, color: Color) {
        self.push(8, (x0, y0, x1, y1), (thickness as f32, 0.0), color, BlendMode::Opaque, 255);
    }
    pub fn draw_rect(&mut self, x0: i32, y0: i32, x1: i32, y1: i32, color: Color) {      // render.rs:941-951
        let (min_x, max_x) = if x0 < x1 { (x0, x1) } else { (x1, x0) };
        let (min_y, max_y) = if y0 < y1 { (y0, y1) } else { (y1, y0) };
        self.draw_line(min_x, min_y, max_x, min_y, color);
        self.draw_line(max_x, min_y, max_x, max_y, color);
        self.draw_line(max_x, max_y, min_x, max_y, color);
        self.draw_line(min_x, max_y, min_x, min_y, color);
    }
    /// Draws the collected lines over the device framebuffer, in the order they were added, and empties the list.
    pub fn flush(&mut self) {
        CTX.with(|&ctx| unsafe { check(ctx, b32_draw_lines(ctx, self.0.as_ptr(), self.0.len() as u32)); });
        self.0.clear();
    }
}

/// Framebuffer::clear_gradient on the device framebuffer.
pub fn clear_gradient(top: Color, bottom: Color) {
    let a = if top.blend == BlendMode::Erase { 0 } else { 255 };
    CTX.with(|&ctx| unsafe { check(ctx, b32_fb_clear_gradient(ctx, top.r, top.g, top.b, bottom.r, bottom.g, bottom.b, a)); });
}

// ---------------------------------------------------------------------------------------------------
// Star pass of Framebuffer::render_skybox (src/rasterizer/render.rs:149-199): the LCG and the libm calls stay
// here, the transform / projection / diamond plotting run on the device.
// ---------------------------------------------------------------------------------------------------
#[repr(C)] #[derive(Clone, Copy)]
pub struct b32_star { dir: [f32; 3], r: u8, g: u8, b: u8, _pad: u8 }
extern "C" {
    fn b32_render_stars(ctx: *mut b32_ctx, stars: *const b32_star, n: u32, cam: *const b32_camera, size: f32) -> c_int;
}

pub fn render_stars(skybox: &crate::world::Skybox, camera: &Camera, time: f32) {
    use std::f32::consts::PI;
    let stars = &skybox.stars;
    let mut rng_seed = stars.seed as u64;
    let mut next_rand = || -> f32 {
        rng_seed = rng_seed.wrapping_mul(1103515245).wrapping_add(12345);
        (rng_seed >> 16) as f32 / 65536.0
    };
    let mut list = Vec::with_capacity(stars.count as usize);
    for _ in 0..stars.count {
        let theta = next_rand() * 2.0 * PI;
        let phi = next_rand() * (skybox.horizon * PI);
        let (y, ring) = (phi.cos(), phi.sin());
        let dir = Vec3::new(ring * theta.cos(), y, ring * theta.sin());
        // the twinkle phase is drawn for visible stars only (render.rs:180-190): same test as the device's
        let cam_z = (dir * 10000.0).dot(camera.basis_z);
        let mut brightness = 1.0f32;
        if cam_z > 0.1 && stars.twinkle_speed > 0.0 {
            let phase = next_rand() * 2.0 * PI;
            brightness = 0.5 + 0.5 * (time * stars.twinkle_speed + phase).sin();
        }
        list.push(b32_star { dir: [dir.x, dir.y, dir.z], r: (stars.color.r as f32 * brightness) as u8,
                             g: (stars.color.g as f32 * brightness) as u8, b: (stars.color.b as f32 * brightness) as u8, _pad: 0 });
    }
    let cam = marshal_camera(camera);
    CTX.with(|&ctx| unsafe { check(ctx, b32_render_stars(ctx, list.as_ptr(), list.len() as u32, &cam, stars.size)); });
}

// ---------------------------------------------------------------------------------------------------
// Placed asset parts: render_asset_parts (src/scene.rs:109-169) with the part resident on the device and the
// per-object rotate-about-Y + translate done there.  `part_mesh` = b32_mesh_upload(part.mesh.to_render_data_textured())
// cached per asset generation; textures as in render_mesh_15 (the part's atlas + CLUT, indexed formats accepted).
// ---------------------------------------------------------------------------------------------------
pub enum b32_mesh {}
#[repr(C)] pub struct b32_placement { facing: f32, cos_f: f32, sin_f: f32, world_pos: [f32; 3] }
extern "C" {
    fn b32_render_mesh_placed(ctx: *mut b32_ctx, mesh: *const b32_mesh, pl: *const b32_placement, cam: *const b32_camera,
                              s: *const b32_settings, fog: *const b32_fog, rgb888: c_int, flags: u32, out: *mut b32_timings) -> c_int;
}

pub unsafe fn render_part_placed(part_mesh: *const b32_mesh, camera: &Camera, render_settings: &RasterSettings, facing: f32,
                                 world_pos: Vec3, fog: Option<(f32, f32, f32, Color)>) -> RasterTimings {
    let pl = b32_placement { facing, cos_f: facing.cos(), sin_f: facing.sin(), world_pos: [world_pos.x, world_pos.y, world_pos.z] };
    let (s, _lights) = marshal_settings(render_settings);
    let cam = marshal_camera(camera);
    let fogc = fog.map(|(start, falloff, cull_distance, c)| b32_fog { start, falloff, cull_distance, r: c.r, g: c.g, b: c.b, blend: blend_u8(c.blend) });
    let mut tm = b32_timings::default();
    CTX.with(|&ctx| check(ctx, b32_render_mesh_placed(ctx, part_mesh, &pl, &cam, &s, fogc.as_ref().map_or(std::ptr::null(), |f| f as *const _),
                                                      (!render_settings.use_rgb555) as c_int, 0, &mut tm)));
    timings(&tm)
}
