"""Writes tests/golden/ref_wasm/*: results of the REFERENCE'S OWN COMPILED rasterizer.

Build container only (needs /root/reference/docs/bonnie-32.wasm).  The binary is interpreted by
oracle/wasm/wasm_interp.cpp; scenes come from tests/refbin_cases.py.  For every scene the fixture stores
sha256 of the inputs, of the RGBA framebuffer and of the z-buffer, and `triangles_drawn`; a reference panic
(trap) is stored as such.  Full frames of a few small scenes are kept in frames.npz for debugging.

    python tests/golden/make_ref_wasm.py [--big]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "wasm"))
import __graft_entry__ as g  # noqa: E402

g.load_package()
from bonnie32_b200 import scenes  # noqa: E402
import refbin_cases  # noqa: E402
from ref_scene import RefRasterizer  # noqa: E402
from ref_wasm import WasmTrap, WASM  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_wasm")
KEEP_FRAMES = {"c1_single_triangle", "c1_single_triangle_float", "c2_1000_tris_64x64_idx8", "mixed_zbuffer", "gouraud_lights"}


def run(R, sc, rgb888=False):
    try:
        rgba, z, drawn = R.render_scene888(sc) if rgb888 else R.render_scene(sc, scenes.expand_texture)
    except WasmTrap as e:
        return {"inputs": refbin_cases.inputs_digest(sc), "trap": str(e)}, None
    a, b = refbin_cases.frame_digest(rgba, z)
    return {"inputs": refbin_cases.inputs_digest(sc), "rgba": a, "z": b, "drawn": int(drawn)}, (rgba, z)


def main():
    os.makedirs(OUT, exist_ok=True)
    big = "--big" in sys.argv
    path = os.path.join(OUT, "render_mesh_15.json")
    res = json.load(open(path)) if os.path.exists(path) else {"scenes": {}}
    res["wasm_sha256"] = hashlib.sha256(open(WASM, "rb").read()).hexdigest()
    res["wasm"] = "docs/bonnie-32.wasm of EBonura/bonnie-32 @ 6ac0a67 (crate version string 0.1.8, rustc 1.92.0)"
    frames = {}
    for sc in refbin_cases.small_scenes() + (refbin_cases.big_scenes() if big else []):
        t = time.time()
        R = RefRasterizer()          # fresh instance per scene: a trap leaves the heap in an unknown state
        rec, fr = run(R, sc)
        res["scenes"][sc.name] = rec
        if fr is not None and sc.name in KEEP_FRAMES:
            frames[sc.name + "/rgba"], frames[sc.name + "/z"] = fr
        print(f"{sc.name:48s} {time.time() - t:6.2f}s {rec.get('drawn', rec.get('trap'))}", flush=True)
    json.dump(res, open(path, "w"), indent=1, sort_keys=True)
    # the RGB888 sibling
    path8 = os.path.join(OUT, "render_mesh.json")
    res8 = {"scenes": {}, "wasm_sha256": res["wasm_sha256"], "wasm": res["wasm"]}
    for sc in refbin_cases.small_scenes888():
        t = time.time()
        rec, fr = run(RefRasterizer(), sc, rgb888=True)
        res8["scenes"][sc.name] = rec
        print(f"{sc.name:48s} {time.time() - t:6.2f}s {rec.get('drawn', rec.get('trap'))}", flush=True)
    json.dump(res8, open(path8, "w"), indent=1, sort_keys=True)
    if frames:
        np.savez_compressed(os.path.join(OUT, "frames.npz"), **frames)


if __name__ == "__main__":
    main()
