"""Writes tests/golden/ref_wasm/skybox.npz + skybox.json: Framebuffer::render_skybox (render.rs:81-299) run by the
reference binary (docs/bonnie-32.wasm, build container only), together with the inputs the C ABI's two skybox entry
points take for the same frame:

  * the sphere mesh `Skybox::generate_mesh` produced inside that call.  generate_mesh is inlined into render_skybox,
    so the mesh is rebuilt here exactly as the decompiled function builds it (33 x 49 vertices at radius 10 000 around
    the camera, two triangles per cell) with the binary's OWN `sinf`, `cosf` and `Skybox::sample_at_direction`
    (separate functions in the binary, called through the interpreter) — no libm of this machine is involved;
  * the star list of render_stars' host half (render.rs:159-196: LCG, spherical direction, visibility, twinkle), computed
    the same way with the binary's libm.

The skybox structs are parsed by the binary itself (`load_level_from_str` on a sample level whose `skybox: None` is
replaced); the level's skybox sits at Level + 12 (recovered from game::renderer's call site).

    python tests/golden/make_ref_wasm_skybox.py
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "wasm"))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from bonnie32_b200 import abi, levels  # noqa: E402
import cases  # noqa: E402
from ref_scene import RefRasterizer  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_wasm")
F_SINF, F_COSF = 2035, 2054
f32 = np.float32

SKIES = {
    "default": b"Some(())",
    "custom": b"Some((zenith_color: (r: 20, g: 40, b: 160), horizon_sky_color: (r: 250, g: 180, b: 90), horizon_ground_color: (r: 90, g: 70, b: 60), "
              b"nadir_color: (r: 10, g: 30, b: 20), horizon: 0.42, horizontal_tint_enabled: true, horizontal_tint_intensity: 0.7, horizontal_tint_spread: 0.9))",
    "stars": b"Some((horizon: 0.6, stars: (enabled: true, count: 150, size: 3.0, twinkle_speed: 1.5, seed: 4242, color: (r: 255, g: 240, b: 200))))",
    "stars_small": b"Some((stars: (enabled: true, count: 400, size: 1.0, twinkle_speed: 0.0, seed: 7)))",
    "stars_mid": b"Some((horizon: 0.8, stars: (enabled: true, count: 90, size: 2.0, twinkle_speed: 0.5, seed: 99)))",
}
# (sky, camera (rot_x, rot_y, position), time, width, height)
CASES = [("default", (0.0, 0.0, (0.0, 0.0, 0.0)), 0.0, 320, 240),
         ("default", (-0.9, 2.2, (1500.0, -300.0, 800.0)), 12.5, 160, 120),
         ("custom", (0.35, 0.7, (100.0, 50.0, -200.0)), 3.25, 320, 240),
         ("custom", (1.2, -2.0, (0.0, 0.0, 0.0)), 100.0, 333, 77),
         ("stars", (-0.6, 0.4, (0.0, 0.0, 0.0)), 7.75, 320, 240),
         ("stars_small", (-1.2, 3.0, (-50.0, 10.0, 5.0)), 1.0, 160, 120),
         ("stars_mid", (-0.7, 1.9, (0.0, 200.0, 0.0)), 42.0, 200, 150)]


def main():
    ref = RefRasterizer()
    w = ref.w
    names = dict((i, n) for i, n in w.mod.find("sinf") + w.mod.find("cosf"))
    assert names.get(F_SINF) == "sinf" and names.get(F_COSF) == "cosf", names
    sinf = lambda x: f32(w.call(F_SINF, float(x)))
    cosf = lambda x: f32(w.call(F_COSF, float(x)))
    txt = levels.brotli_decompress(open("/root/reference/assets/samples/levels/West.ron", "rb").read())
    sky_ptr = {}
    for name, ron in SKIES.items():
        t = txt.replace(b"skybox: None", b"skybox: " + ron)
        p = w.put(t, 1)
        out = w.alloc(1024, 8)
        w.call("load_level_from_str", out, p, len(t))
        assert w.read(out + 258, 1)[0] != 2, "skybox did not parse"
        sky_ptr[name] = out + 12
    arrays, meta = {}, {"binary": "docs/bonnie-32.wasm (crate 0.1.8)", "cases": {}}
    for k, (sky, (rx, ry, pos), time, width, height) in enumerate(CASES):
        cam = cases._rotated_camera(rx, ry, pos)
        sp = sky_ptr[sky]
        ref._allocs = []
        cam_ptr = ref._camera(cam)
        # ---- the binary's frame
        fb = ref.new_framebuffer(width, height)
        w.call("render_skybox", fb, sp, cam_ptr, float(time))
        frame, _ = ref.read_framebuffer(fb)
        ref.free_framebuffer(fb)
        # ---- generate_mesh as the binary runs it (decompiled render_skybox, first two loops)
        cpos = [f32(x) for x in cam.position]
        verts = np.zeros(33 * 49, dtype=abi.SKY_VERTEX_DTYPE)
        sin_t = [sinf((f32(j) * f32(6.2831854820251465)) / f32(48.0)) for j in range(49)]
        cos_t = [cosf((f32(j) * f32(6.2831854820251465)) / f32(48.0)) for j in range(49)]
        for i in range(33):
            phi = (f32(i) * f32(3.1415927410125732)) * f32(0.03125)
            y = cosf(phi) * f32(10000.0) + cpos[1]
            s_phi = sinf(phi)
            for j in range(49):
                theta = (f32(j) * f32(6.2831854820251465)) / f32(48.0)
                z = (s_phi * sin_t[j]) * f32(10000.0) + cpos[2]
                x = (s_phi * cos_t[j]) * f32(10000.0) + cpos[0]
                c = w.call("sample_at_direction", sp, float(theta), float(phi), float(time))
                v = verts[i * 49 + j]
                v["pos"] = (x, y, z)
                v["rgb"] = ((c >> 8) & 255, (c >> 16) & 255, (c >> 24) & 255)
        faces = []
        for r in range(32):
            for c_ in range(48):
                bl = (r + 1) * 49 + c_
                faces += [(bl - 49, bl, bl - 48), (bl - 48, bl, bl + 1)]
        faces = np.array(faces, np.uint32)
        # ---- render_stars' host half (render.rs:159-196) with the binary's libm
        raw = w.read(sp, 248)
        horizon = np.frombuffer(raw[128:132], "<f4")[0]
        st_rgb = raw[225:228]
        st_size, st_twinkle = np.frombuffer(raw[228:236], "<f4")
        st_seed, = struct.unpack("<I", raw[236:240])
        st_count, = struct.unpack("<H", raw[240:242])
        st_enabled = raw[242]
        stars = []
        if st_enabled:
            seed = int(st_seed)
            def next_rand():
                nonlocal seed
                seed = (seed * 1103515245 + 12345) & 0xFFFFFFFFFFFFFFFF
                return f32(seed >> 16) / f32(65536.0)
            PI = f32(3.1415927410125732)
            bx, by, bz = ([f32(x) for x in b] for b in (cam.basis_x, cam.basis_y, cam.basis_z))
            for _ in range(st_count):
                theta = next_rand() * f32(2.0) * PI
                phi_max = horizon * PI
                phi = next_rand() * phi_max
                y = cosf(phi); ring = sinf(phi)
                x = ring * cosf(theta); z = ring * sinf(theta)
                d = (x * f32(10000.0), y * f32(10000.0), z * f32(10000.0))
                cz = d[0] * bz[0] + d[1] * bz[1] + d[2] * bz[2]                 # Vec3::dot, math.rs:23-25
                if cz > f32(0.1):
                    brightness = f32(1.0)
                    if st_twinkle > 0.0:
                        phase = next_rand() * f32(2.0) * PI
                        brightness = f32(0.5) + f32(0.5) * sinf(f32(time) * st_twinkle + phase)
                    col = tuple(int(np.clip(np.trunc(f32(ch) * brightness), 0, 255)) for ch in st_rgb)       # `as u8` saturates
                else:
                    col = (0, 0, 0)                                             # never drawn: the device rejects it by the same test
                stars.append(((x, y, z), col, 0))
        stars = np.array(stars, dtype=abi.STAR_DTYPE) if stars else np.zeros(0, abi.STAR_DTYPE)
        key = f"case{k}"
        arrays[key + "_verts"], arrays[key + "_faces"], arrays[key + "_stars"] = verts, faces, stars
        if width * height <= 160 * 120:
            arrays[key + "_frame"] = frame                                      # small frames are kept whole (debugging aid)
        meta["cases"][key] = {"sky": sky, "camera": [rx, ry, list(pos)], "time": time, "width": width, "height": height,
                              "star_size": float(st_size) if st_enabled else 0.0, "n_stars": int(len(stars)),
                              "rgba": hashlib.sha256(frame.tobytes()).hexdigest(), "nonzero_pixels": int((frame[..., 3] != 0).sum())}
        print(key, sky, width, height, "stars", len(stars), "drawn pixels", meta["cases"][key]["nonzero_pixels"])
    np.savez_compressed(os.path.join(OUT, "skybox.npz"), **arrays)
    json.dump(meta, open(os.path.join(OUT, "skybox.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
