// b32_raster.hpp — header-only C++ mirror of the reference's rasterizer interface over the C ABI.
//
// Same names and argument meaning as /root/reference/src/rasterizer: Framebuffer (render.rs:10-45),
// render_mesh_15 (render.rs:2302-2310), RasterTimings (types.rs:1499-1514).  Errors that are panics
// in the reference (bad vertex index, NaN sort key) are thrown as b32::Error.  No CPU fallback.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "b32_raster.h"

namespace b32 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("b32 error " + std::to_string(c) + ": " + m), code(c) {}
};

class Context {
public:
    explicit Context(int device = 0) {
        int rc = b32_ctx_create(device, &ctx_);
        if (rc != B32_OK) throw Error(rc, "b32_ctx_create failed (no CUDA device; there is no CPU fallback)");
    }
    ~Context() { b32_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    b32_ctx* get() const { return ctx_; }
    void check(int rc) const { if (rc != B32_OK) throw Error(rc, b32_last_error(ctx_)); }
    void set_textures(const std::vector<b32_tex_desc>& t) { check(b32_textures_set(ctx_, t.data(), (uint32_t)t.size())); }
    void set_textures_rgb888(const std::vector<b32_tex8_desc>& t) { check(b32_textures_set_rgb888(ctx_, t.data(), (uint32_t)t.size())); }
    void sync() { check(b32_sync(ctx_)); }             // errors of enqueue-only calls surface here
private:
    b32_ctx* ctx_ = nullptr;
};

// struct Framebuffer, render.rs:10-45: pixels (RGBA8) + zbuffer (f32) live on the device.
class Framebuffer {
public:
    Framebuffer(Context& c, uint32_t w, uint32_t h) : c_(c) { resize(w, h); clear_transparent(); }
    void resize(uint32_t w, uint32_t h) { c_.check(b32_fb_resize(c_.get(), w, h)); width = w; height = h; }     // :27-34
    void clear(uint8_t r, uint8_t g, uint8_t b, bool erase = false) { c_.check(b32_fb_clear(c_.get(), r, g, b, erase ? 0 : 255)); }  // :36-45
    void clear_transparent() { c_.check(b32_fb_clear(c_.get(), 0, 0, 0, 0)); }                                    // :48-56
    void clear_gradient(const uint8_t top[3], const uint8_t bottom[3], bool top_erase = false) {                    // :60-77
        c_.check(b32_fb_clear_gradient(c_.get(), top[0], top[1], top[2], bottom[0], bottom[1], bottom[2], top_erase ? 0 : 255));
    }
    // render_skybox (:81-146): the sphere pass on the mesh Skybox::generate_mesh returned, then the stars the host prepared
    void render_skybox_mesh(const std::vector<b32_sky_vertex>& v, const std::vector<uint32_t>& faces, const b32_camera& cam) {
        c_.check(b32_render_skybox_mesh(c_.get(), v.data(), (uint32_t)v.size(), faces.data(), (uint32_t)(faces.size() / 3), &cam));
    }
    void render_stars(const std::vector<b32_star>& stars, const b32_camera& cam, float size) {
        c_.check(b32_render_stars(c_.get(), stars.data(), (uint32_t)stars.size(), &cam, size));
    }
    // draw_line*, draw_circle*, draw_thick_line, draw_filled_rect (:631-972): a list drawn with the result of the calls made in order
    void draw_lines(const std::vector<b32_line>& lines) { c_.check(b32_draw_lines(c_.get(), lines.data(), (uint32_t)lines.size())); }
    void upload(const uint8_t* rgba, const float* z) { c_.check(b32_fb_upload(c_.get(), rgba, z)); }
    std::vector<uint8_t> pixels() { std::vector<uint8_t> p((size_t)width * height * 4); c_.check(b32_fb_download(c_.get(), p.data(), nullptr)); return p; }
    std::vector<float> zbuffer() { std::vector<float> z((size_t)width * height); c_.check(b32_fb_download(c_.get(), nullptr, z.data())); return z; }
    uint32_t width = 0, height = 0;
    Context& context() { return c_; }
private:
    Context& c_;
};

using RasterTimings = b32_timings;

// render_mesh_15, render.rs:2302-2310 (textures are set on the context: Context::set_textures)
inline RasterTimings render_mesh_15(Framebuffer& fb, const std::vector<b32_vertex>& vertices, const std::vector<b32_face>& faces,
                                    const b32_camera& camera, const b32_settings& settings, const b32_fog* fog = nullptr) {
    RasterTimings tm{};
    fb.context().check(b32_render_mesh_15(fb.context().get(), vertices.data(), (uint32_t)vertices.size(), faces.data(),
                                          (uint32_t)faces.size(), &camera, &settings, fog, &tm));
    return tm;
}

// render_mesh (RGB888 sibling), render.rs:1971-1978 (textures: Context::set_textures_rgb888)
inline RasterTimings render_mesh(Framebuffer& fb, const std::vector<b32_vertex>& vertices, const std::vector<b32_face>& faces,
                                 const b32_camera& camera, const b32_settings& settings) {
    RasterTimings tm{};
    fb.context().check(b32_render_mesh(fb.context().get(), vertices.data(), (uint32_t)vertices.size(), faces.data(),
                                       (uint32_t)faces.size(), &camera, &settings, &tm));
    return tm;
}

// Device-resident geometry (static rooms, asset parts): upload once, render per frame.
class Mesh {
public:
    Mesh(Context& c, const std::vector<b32_vertex>& vertices, const std::vector<b32_face>& faces) : c_(c) {
        c_.check(b32_mesh_upload(c_.get(), vertices.data(), (uint32_t)vertices.size(), faces.data(), (uint32_t)faces.size(), &m_));
    }
    ~Mesh() { b32_mesh_free(c_.get(), m_); }
    Mesh(const Mesh&) = delete;
    Mesh& operator=(const Mesh&) = delete;
    RasterTimings render_15(const b32_camera& camera, const b32_settings& settings, const b32_fog* fog = nullptr) {
        RasterTimings tm{};
        c_.check(b32_render_mesh_15_resident(c_.get(), m_, &camera, &settings, fog, &tm));
        return tm;
    }
    // Framebuffer::clear(clear_rgba) + render_mesh_15 as one enqueued frame (no wait; Context::sync / a download completes it)
    void frame_15_enqueue(const uint8_t* clear_rgba, const b32_camera& camera, const b32_settings& settings, const b32_fog* fog = nullptr) {
        c_.check(b32_frame_15_enqueue(c_.get(), clear_rgba, m_, &camera, &settings, fog));
    }
    // one part of render_asset_parts (src/scene.rs:109-169): rotate about Y by `facing`, translate, render
    RasterTimings render_placed(float facing, float cos_f, float sin_f, const float world_pos[3], const b32_camera& camera,
                                const b32_settings& settings, const b32_fog* fog = nullptr, bool rgb888 = false) {
        b32_placement pl{facing, cos_f, sin_f, {world_pos[0], world_pos[1], world_pos[2]}};
        RasterTimings tm{};
        c_.check(b32_render_mesh_placed(c_.get(), m_, &pl, &camera, &settings, fog, rgb888 ? 1 : 0, 0u, &tm));
        return tm;
    }
private:
    Context& c_;
    b32_mesh* m_ = nullptr;
};

}  // namespace b32
