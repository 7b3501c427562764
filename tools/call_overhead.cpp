// call_overhead.cpp — what one render_mesh_15 call costs a native (C/C++/Rust) host, per mesh size (GPU box only).
// build: g++ -O2 -std=c++17 -Iinclude tools/call_overhead.cpp -Lbonnie-32_b200 -lb32raster -Wl,-rpath,'$ORIGIN/../bonnie-32_b200' -o build/call_overhead
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "b32_raster.h"

static uint64_t sm_state = 0xB3200002ull;
static double u01() {
    sm_state += 0x9E3779B97F4A7C15ull;
    uint64_t z = sm_state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (double)(z >> 40) / 16777216.0;
}

int main() {
    b32_ctx* ctx = nullptr;
    if (b32_ctx_create(0, &ctx) != B32_OK) { fprintf(stderr, "no CUDA device\n"); return 1; }
    b32_fb_resize(ctx, 320, 240);
    std::vector<uint16_t> tex(64 * 64);
    for (auto& t : tex) t = (uint16_t)(u01() * 32768.0) | 1;
    b32_tex_desc td{64, 64, B32_TEX_RGB555, B32_BLEND_OPAQUE, tex.data(), nullptr, 0};
    b32_textures_set(ctx, &td, 1);
    b32_camera cam{{0, 0, 0}, {-1, 0, 0}, {0, -1, 0}, {0, 0, 1}};
    b32_settings st{};
    st.affine_textures = 1; st.backface_cull = 1; st.dithering = 1; st.use_rgb555 = 1; st.use_fixed_point = 1; st.ambient = 0.3f;
    for (uint32_t nf : {1u, 100u, 1000u, 10000u, 100000u}) {
        std::vector<b32_vertex> v(nf * 3); std::vector<b32_face> f(nf);
        for (uint32_t t = 0; t < nf; ++t) {
            double cz = 2.0 + 58.0 * u01(), cx = (2 * u01() - 1) * 0.5 * (cz + 5), cy = (2 * u01() - 1) * 0.4 * (cz + 5), r = 0.04 * (cz + 5);
            for (int k = 0; k < 3; ++k) {
                b32_vertex& q = v[t * 3 + k];
                q.pos[0] = (float)(cx + r * (2 * u01() - 1)); q.pos[1] = (float)(cy + r * (2 * u01() - 1)); q.pos[2] = (float)(cz + r * (2 * u01() - 1));
                q.uv[0] = (float)(2 * u01()); q.uv[1] = (float)(2 * u01()); q.normal[0] = q.normal[1] = 0; q.normal[2] = -1;
                q.r = q.g = q.b = 128; q.blend = 0;
            }
            f[t] = b32_face{t * 3, t * 3 + 1, t * 3 + 2, B32_FACE_FLAGS(0, 0, 1, 255)};
        }
        b32_mesh* mesh = nullptr;
        b32_mesh_upload(ctx, v.data(), (uint32_t)v.size(), f.data(), nf, &mesh);
        b32_timings tm{};
        const int reps = 300;
        auto run = [&](const char* what, auto call) {
            for (int i = 0; i < 20; ++i) call();
            b32_sync(ctx);
            auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < reps; ++i) call();
            b32_sync(ctx);
            double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
            printf("nf=%6u %-44s %8.1f us/call\n", nf, what, us);
        };
        run("render_mesh_15_resident (blocking, timings)", [&] { b32_render_mesh_15_resident(ctx, mesh, &cam, &st, nullptr, &tm); });
        run("render_mesh_15 host buffers (blocking)", [&] { b32_render_mesh_15(ctx, v.data(), (uint32_t)v.size(), f.data(), nf, &cam, &st, nullptr, &tm); });
        run("render_mesh_15_enqueue (no wait)", [&] { b32_render_mesh_15_enqueue(ctx, mesh, &cam, &st, nullptr); });
        uint8_t clear[4] = {20, 22, 28, 255};
        run("frame_15_enqueue (clear + render, graph)", [&] { b32_frame_15_enqueue(ctx, clear, mesh, &cam, &st, nullptr); });
        b32_mesh_free(ctx, mesh);
    }
    b32_ctx_destroy(ctx);
    return 0;
}
