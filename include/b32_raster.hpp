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

}  // namespace b32
