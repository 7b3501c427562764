"""Load/save of the BASELINE config 3 fixtures (sample world scenes) — see tests/golden/make_c3.py."""
import glob
import os

import numpy as np

from bonnie32_b200 import abi, levels
from bonnie32_b200.raster import Camera, Texture15

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# RasterSettings::default() minus backface_wireframe (z-buffer ON, Gouraud, dither), and the painter's variant
MODES = {"zbuffer": dict(), "painter": dict(use_zbuffer=False)}


def save_scene(sc, path):
    d = {"n_rooms": np.int64(len(sc.rooms)), "n_tex": np.int64(len(sc.textures)),
         "cam": np.concatenate([sc.camera.position, sc.camera.basis_x, sc.camera.basis_y, sc.camera.basis_z]).astype(np.float32)}
    for i, rc in enumerate(sc.rooms):
        d[f"v{i}"] = rc.vertices.view(np.uint8)
        d[f"f{i}"] = rc.faces.view(np.uint8)
        fog = rc.fog
        d[f"m{i}"] = np.array([rc.ambient, 1.0 if fog else 0.0] + (list(fog[:3]) + list(fog[3]) if fog else [0.0] * 6), dtype=np.float64)
    for i, t in enumerate(sc.textures):
        d[f"t{i}"] = np.asarray(t.pixels, dtype=np.uint16).reshape(t.height, t.width)
    # placed asset parts (scene.rs:219-259) and the point lights collected from placed assets (scene.rs:32-70)
    d["n_parts"] = np.int64(len(sc.parts))
    for i, pc in enumerate(sc.parts):
        d[f"pv{i}"] = pc.vertices.view(np.uint8)
        d[f"pf{i}"] = pc.faces.view(np.uint8)
        fog = pc.fog
        d[f"pm{i}"] = np.array([pc.facing, *pc.world_pos, 1.0 if pc.double_sided else 0.0, pc.ambient, 1.0 if fog else 0.0]
                               + (list(fog[:3]) + list(fog[3]) if fog else [0.0] * 6), dtype=np.float64)
    d["lights"] = np.array([[*l.position, l.radius, l.intensity, *l.color] for l in sc.lights], dtype=np.float64).reshape(-1, 8)
    np.savez_compressed(path, **d)


def load_scene(path):
    z = np.load(path)
    cam = Camera()
    c = z["cam"]
    cam.position, cam.basis_x, cam.basis_y, cam.basis_z = (c[0:3].copy(), c[3:6].copy(), c[6:9].copy(), c[9:12].copy())
    rooms = []
    for i in range(int(z["n_rooms"])):
        m = z[f"m{i}"]
        fog = (float(m[2]), float(m[3]), float(m[4]), (int(m[5]), int(m[6]), int(m[7]))) if m[1] else None
        rooms.append(levels.RoomCall(z[f"v{i}"].view(abi.VERTEX_DTYPE).copy(), z[f"f{i}"].view(abi.FACE_DTYPE).copy(), float(np.float32(m[0])), fog))
    texs = []
    for i in range(int(z["n_tex"])):
        a = z[f"t{i}"]
        texs.append(Texture15(a.shape[1], a.shape[0], a.reshape(-1).copy()))
    name = os.path.basename(path)[3:-4]
    sc = levels.LevelScene(name, rooms, texs, cam)
    for i in range(int(z["n_parts"]) if "n_parts" in z else 0):
        m = z[f"pm{i}"]
        fog = (float(m[7]), float(m[8]), float(m[9]), (int(m[10]), int(m[11]), int(m[12]))) if m[6] else None
        sc.parts.append(levels.PartCall(z[f"pv{i}"].view(abi.VERTEX_DTYPE).copy(), z[f"pf{i}"].view(abi.FACE_DTYPE).copy(), float(np.float32(m[0])),
                                        tuple(float(np.float32(x)) for x in m[1:4]), bool(m[4]), float(np.float32(m[5])), fog))
    from bonnie32_b200.raster import Light
    for row in (z["lights"] if "lights" in z else []):
        l = Light.point(np.asarray(row[0:3], np.float32), float(np.float32(row[3])), float(np.float32(row[4])))
        l.color = tuple(int(x) for x in row[5:8])
        sc.lights.append(l)
    return sc


def scene_paths():
    return sorted(glob.glob(os.path.join(GOLDEN, "c3_*.npz")))
