"""Writes tests/golden/ref_wasm/prims.json: lists of overlay primitives drawn by the reference binary's own
Framebuffer::draw_line_3d_impl / draw_circle / draw_thick_line (docs/bonnie-32.wasm, build container only), one call
per primitive in list order, over tests/refbin_prims.py's backgrounds; kept: sha256 of the final pixels.

    python tests/golden/make_ref_wasm_prims.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "wasm"))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from bonnie32_b200 import abi  # noqa: E402
from ref_scene import RefRasterizer  # noqa: E402
import refbin_prims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_wasm", "prims.json")


def color_word(l):
    """struct Color in the wasm32 build: blend @0, r @1, g @2, b @3 (passed by value as one i32)."""
    return int(l["blend"]) | (int(l["rgb"][0]) << 8) | (int(l["rgb"][1]) << 16) | (int(l["rgb"][2]) << 24)


def draw(ref, fb, l):
    w = ref.w
    k = int(l["kind"])
    if k in (abi.LINE_3D, abi.LINE_3D_OVERLAY):
        w.call("draw_line_3d_impl", fb, int(l["x0"]), int(l["y0"]), float(l["z0"]), int(l["x1"]), int(l["y1"]), float(l["z1"]),
               color_word(l), 1 if k == abi.LINE_3D_OVERLAY else 0)
    elif k == abi.LINE_CIRCLE:
        w.call("Framebuffer11draw_circle", fb, int(l["x0"]), int(l["y0"]), int(l["x1"]), color_word(l))
    elif k == abi.LINE_THICK:
        w.call("draw_thick_line", fb, int(l["x0"]), int(l["y0"]), int(l["x1"]), int(l["y1"]), int(l["z0"]), color_word(l))
    else:
        raise ValueError(k)


def main():
    ref = RefRasterizer()
    out = {"binary": "docs/bonnie-32.wasm (crate 0.1.8)", "cases": {}}
    for name, w, h, seed, n in refbin_prims.CASES:
        rgba, z = refbin_prims.background(w, h, seed)
        lines = refbin_prims.primitives(w, h, seed, n)
        fb = ref.new_framebuffer(w, h, rgba, z)
        for l in lines:
            draw(ref, fb, l)
        got, got_z = ref.read_framebuffer(fb)
        ref.free_framebuffer(fb)
        assert np.array_equal(got_z.view(np.uint32), z.view(np.uint32))           # these primitives never write depth
        changed = int((got != rgba).any(-1).sum())
        out["cases"][name] = {"inputs": hashlib.sha256(rgba.tobytes() + z.tobytes() + lines.tobytes()).hexdigest(),
                              "rgba": hashlib.sha256(got.tobytes()).hexdigest(), "pixels_changed": changed}
        print(name, changed, "pixels changed")
    # ---- draw_line through draw::draw_3d_line_clipped(fb, &camera, &p0, &p1, color)
    from bonnie32_b200.raster import Camera
    out["clipped"] = {}
    for name, w, h, seed, n in refbin_prims.CLIPPED:
        rgba, z = refbin_prims.background(w, h, seed)
        p0, p1, rgb, ends = refbin_prims.clipped_segments(w, h, seed, n)
        fb = ref.new_framebuffer(w, h, rgba, z)
        ref._allocs = []
        cam = ref._camera(Camera())
        a, b = ref.w.alloc(12, 4), ref.w.alloc(12, 4)
        for i in range(n):
            ref.w.write(a, p0[i].tobytes()); ref.w.write(b, p1[i].tobytes())
            ref.w.call("draw_3d_line_clipped", fb, cam, a, b, (int(rgb[i, 0]) << 8) | (int(rgb[i, 1]) << 16) | (int(rgb[i, 2]) << 24))
        got, _ = ref.read_framebuffer(fb)
        ref.free_framebuffer(fb)
        out["clipped"][name] = {"inputs": hashlib.sha256(rgba.tobytes() + p0.tobytes() + p1.tobytes() + rgb.tobytes()).hexdigest(),
                                "rgba": hashlib.sha256(got.tobytes()).hexdigest(), "pixels_changed": int((got != rgba).any(-1).sum())}
        print(name, out["clipped"][name]["pixels_changed"], "pixels changed")
    json.dump(out, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
