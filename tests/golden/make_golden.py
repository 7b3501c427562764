"""Regenerates tests/golden/ with the numpy model (oracle/pymodel.py), the restatement that is
independent of the C++ oracle.  Run from the repo root:  python tests/golden/make_golden.py

hashes.json : sha256 of the RGBA8 framebuffer bytes and of the f32 z-buffer bytes per scene
*.npz       : full framebuffers of the small BASELINE configs (C1, C2)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
from oracle import pymodel  # noqa: E402
import cases  # noqa: E402


def render(sc):
    rgba, z = pymodel.fb_clear(sc.width, sc.height, sc.clear)
    order = pymodel.render_mesh_15(rgba, z, sc.vertices, sc.faces, sc.textures, sc.camera, sc.settings, sc.fog)
    return rgba, z, order


def digest(rgba, z, order):
    return {"rgba_sha256": hashlib.sha256(np.ascontiguousarray(rgba).tobytes()).hexdigest(),
            "z_sha256": hashlib.sha256(np.ascontiguousarray(z).tobytes()).hexdigest(),
            "triangles_drawn": len(order),
            "order_sha256": hashlib.sha256(np.asarray(order, dtype=np.uint32).tobytes()).hexdigest()}


def golden_scenes(full=True):
    s = pkg.scenes
    out = [s.scene_c1(True), s.scene_c1(False), s.scene_c2(), s.scene_c2(use_zbuffer=True)]
    out[-1].name += "_zbuffer"
    out += cases.feature_scenes() + [cases.big_triangle_scene()]
    if full:
        c4z = s.scene_c4(use_zbuffer=True); c4z.name += "_zbuffer"
        out += [s.scene_c4(), c4z] + [s.scene_c5(k) for k in range(8)]
    return out


def main_rgb888():
    """hashes_rgb888.json: the RGB888 sibling (render_mesh) on cases.rgb888_scenes(), numpy model."""
    hashes = {}
    for sc in cases.rgb888_scenes():
        rgba, z = pymodel.fb_clear(sc.width, sc.height, sc.clear)
        order = pymodel.render_mesh(rgba, z, sc.vertices, sc.faces, sc.textures8, sc.camera, sc.settings)
        hashes[sc.name] = digest(rgba, z, order)
        print(f"{sc.name:45s} drawn={len(order)}")
    with open(os.path.join(HERE, "hashes_rgb888.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)


def main_sky():
    """hashes_sky.json: the skybox sphere pass on cases.sky_cases(), numpy model."""
    hashes = {}
    for name, w, h, cam in cases.sky_cases():
        sv, f = cases.sky_mesh(cam.position)
        rgba, _ = pymodel.fb_clear(w, h, (0, 0, 0))
        pymodel.render_skybox_mesh(rgba, sv, f, cam)
        hashes[name] = hashlib.sha256(rgba.tobytes()).hexdigest()
        print(name, hashes[name][:16])
    with open(os.path.join(HERE, "hashes_sky.json"), "w") as f_:
        json.dump(hashes, f_, indent=1, sort_keys=True)


def main():
    if "--rgb888" in sys.argv:
        return main_rgb888()
    if "--sky" in sys.argv:
        return main_sky()
    hashes = {}
    for sc in golden_scenes():
        t = time.time()
        rgba, z, order = render(sc)
        hashes[sc.name] = digest(rgba, z, order)
        print(f"{sc.name:45s} {time.time() - t:6.1f}s drawn={len(order)}")
        if sc.name.startswith(("c1_", "c2_")):
            np.savez_compressed(os.path.join(HERE, sc.name + ".npz"), rgba=rgba, z=z, order=np.asarray(order, dtype=np.uint32))
    with open(os.path.join(HERE, "hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
