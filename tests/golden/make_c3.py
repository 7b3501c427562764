"""BASELINE config 3 fixtures: the sample world scenes of the reference (assets/samples/levels/*.ron)
assembled into render_mesh_15 arguments by bonnie-32_b200/levels.py, and their golden framebuffers
rendered by the numpy model (oracle/pymodel.py).

Run in the container that has /root/reference:   python tests/golden/make_c3.py
Writes tests/golden/c3_<level>.npz (geometry per room, the textures the level uses, camera, per-room
ambient/fog) and the hashes into tests/golden/c3_hashes.json.  The sample levels and texture packs are
the reference's own sample assets (CC0 / free-to-use packs, /root/reference/THIRD_PARTY.md); only
what a level references is stored, quantised to RGB555.
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
from bonnie32_b200 import levels  # noqa: E402
from oracle import pymodel  # noqa: E402
import c3  # noqa: E402

REF = "/root/reference/assets/samples"


def place(pc):
    """render_asset_parts' vertex transform (scene.rs:121-160), float32, in the reference's operation order."""
    from bonnie32_b200.raster import libm_cosf, libm_sinf
    f32 = np.float32
    facing, wp = f32(pc.facing), [f32(x) for x in pc.world_pos]
    if not (abs(facing) > f32(0.0001) or any(abs(x) > f32(0.0001) for x in wp)):
        return pc.vertices
    c, s = f32(libm_cosf(float(facing))), f32(libm_sinf(float(facing)))
    v = pc.vertices.copy()
    x, y, z = pc.vertices["pos"][:, 0], pc.vertices["pos"][:, 1], pc.vertices["pos"][:, 2]
    v["pos"][:, 0] = (x * c - z * s) + wp[0]
    v["pos"][:, 1] = y + wp[1]
    v["pos"][:, 2] = (x * s + z * c) + wp[2]
    nx, nz = pc.vertices["normal"][:, 0], pc.vertices["normal"][:, 2]
    v["normal"][:, 0] = nx * c - nz * s
    v["normal"][:, 2] = nx * s + nz * c
    return v


def main():
    packs = levels.load_texture_packs(os.path.join(REF, "texture-packs"))
    print(len(packs), "textures in", REF)
    hashes = {}
    for fn in sorted(os.listdir(os.path.join(REF, "levels"))):
        if not fn.endswith(".ron"):
            continue
        sc = levels.assemble_level(os.path.join(REF, "levels", fn), packs, assets_dir=os.path.join(REF, "assets"),
                                   user_textures_dir=os.path.join(REF, "textures"))
        c3.save_scene(sc, os.path.join(HERE, f"c3_{sc.name}.npz"))
        sc = c3.load_scene(os.path.join(HERE, f"c3_{sc.name}.npz"))       # golden is made from what is stored
        for mode, kw in c3.MODES.items():
            t = time.time()
            rgba, z = pymodel.fb_clear(sc.width, sc.height, sc.clear)
            drawn = 0
            for rc in sc.rooms:
                order = pymodel.render_mesh_15(rgba, z, rc.vertices, rc.faces, sc.textures, sc.camera, sc.settings(rc.ambient, **kw), rc.fog)
                drawn += len(order)
            for pc in sc.parts:                          # render_asset_parts, scene.rs:109-169
                order = pymodel.render_mesh_15(rgba, z, place(pc), pc.faces, sc.textures, sc.camera, sc.part_settings(pc, **kw), pc.fog)
                drawn += len(order)
            hashes[f"{sc.name}:{mode}"] = {"rgba_sha256": hashlib.sha256(rgba.tobytes()).hexdigest(),
                                           "z_sha256": hashlib.sha256(z.tobytes()).hexdigest(), "triangles_drawn": drawn}
            cov = int((rgba[..., :3] != np.array(sc.clear, np.uint8)).any(-1).sum())
            print(f"{sc.name:12s} {mode:8s} rooms={len(sc.rooms)} parts={len(sc.parts)} lights={len(sc.lights)} tris={sum(len(r.faces) for r in sc.rooms)} drawn={drawn} "
                  f"covered_px={cov} textures={len(sc.textures)} {time.time() - t:.1f}s")
    json.dump(hashes, open(os.path.join(HERE, "c3_hashes.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
