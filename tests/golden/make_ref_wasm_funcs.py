"""Writes tests/golden/ref_wasm/functions.npz: direct calls of two named functions of the reference binary
(docs/bonnie-32.wasm, build container only):

  fixed::project_fixed            (fixed.rs:424-441)   24 000 vertices x 8 cameras x 3 framebuffer sizes, incl. non-finite,
                                                       huge, denormal and |denom| < 256 inputs
  render::shade_multi_light_color (render.rs:1013-1071) 6 000 (normal, position) pairs x 6 light sets (Directional, Point,
                                                       coloured, disabled, zero radius, degenerate distances)

    python tests/golden/make_ref_wasm_funcs.py
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "wasm"))
import __graft_entry__ as g  # noqa: E402

g.load_package()
import cases  # noqa: E402
from ref_wasm import RefWasm  # noqa: E402
import refbin_funcs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_wasm", "functions.npz")


def main():
    w = RefWasm()
    res = {}
    # ---- project_fixed(sret, &world, &cam_pos, &bx, &by, &bz, width, height) -> (i32 sx, i32 sy, f32 depth)
    world, cam_idx, size_idx = refbin_funcs.project_inputs()
    cams = refbin_funcs.cameras()
    out = w.alloc(16, 4)
    wp = w.alloc(12, 4)
    cam_ptrs = []
    for c in cams:
        cam_ptrs.append([w.put(np.asarray(a, np.float32), 4) for a in (c.position, c.basis_x, c.basis_y, c.basis_z)])
    sx = np.empty(len(world), np.int32); sy = np.empty(len(world), np.int32); dz = np.empty(len(world), np.float32)
    for i in range(len(world)):
        w.write(wp, world[i].tobytes())
        cp = cam_ptrs[cam_idx[i]]
        wd, ht = refbin_funcs.SIZES[size_idx[i]]
        w.call('project_fixed', out, wp, cp[0], cp[1], cp[2], cp[3], wd, ht)
        sx[i], sy[i], bits = struct.unpack('<iiI', w.read(out, 12))
        dz[i] = np.frombuffer(struct.pack('<I', bits), np.float32)[0]
    res["project_sx"], res["project_sy"], res["project_depth"] = sx, sy, dz
    # ---- shade_multi_light_color(sret, &normal, &world_pos, lights.ptr, lights.len, ambient) -> (f32, f32, f32)
    normal, pos, set_idx, ambient = refbin_funcs.shade_inputs()
    sets = refbin_funcs.light_sets()
    set_ptrs = []
    for ls in sets:
        rec = b''
        for l in ls:
            name = w.put(b'L')
            payload = [0.0] * 8
            if int(l.type) == 0:
                payload[0:3] = [float(x) for x in l.direction]
            else:
                payload[0:3] = [float(x) for x in l.position]; payload[3] = float(l.radius)
            rec += struct.pack('<I8f', int(l.type), *payload) + struct.pack('<III', 1, name, 1) \
                + bytes([0, l.color[0], l.color[1], l.color[2]]) + struct.pack('<f', l.intensity) + bytes([1 if l.enabled else 0, 0, 0, 0])
        set_ptrs.append((w.put(rec, 4) if rec else 4, len(ls)))
    npn = w.alloc(12, 4); npp = w.alloc(12, 4)
    shade = np.empty((len(normal), 3), np.float32)
    for i in range(len(normal)):
        w.write(npn, normal[i].tobytes()); w.write(npp, pos[i].tobytes())
        lp, ln = set_ptrs[set_idx[i]]
        w.call('shade_multi_light_color', out, npn, npp, lp, ln, float(ambient[i]))
        shade[i] = np.frombuffer(w.read(out, 12), np.float32)
    res["shade"] = shade
    # ---- Framebuffer::set_pixel_blended_15(fb, x, y, color15, mode) (render.rs:475-501): one pixel per input, back colour preloaded
    c15, back, mode = refbin_funcs.blend_inputs()
    n = len(c15)
    fb = w.alloc(32, 4)
    w.call('Framebuffer3new', fb, n, 1)
    pix_ptr = struct.unpack('<I', w.read(fb + 4, 4))[0]
    px = np.zeros((n, 4), np.uint8); px[:, :3] = back; px[:, 3] = 77
    w.write(pix_ptr, px)
    for i in range(n):
        w.call('set_pixel_blended_15', fb, i, 0, int(c15[i]), int(mode[i]))
    res["blend15"] = np.frombuffer(w.read(pix_ptr, n * 4), np.uint8).reshape(n, 4).copy()
    # ---- Framebuffer::clear(fb, color) (render.rs:36-45)
    clears = []
    for word in (0x1C161400, 0xFF000005, 0x01020302, 0x00000000):          # blend @0, r @1, g @2, b @3
        w.call('Framebuffer5clear', fb, word)
        zp = struct.unpack('<I', w.read(fb + 16, 4))[0]
        clears.append(np.concatenate([np.frombuffer(w.read(pix_ptr, 8), np.uint8), np.frombuffer(w.read(zp, 8), np.uint8)]))
    res["clear"] = np.stack(clears)
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
