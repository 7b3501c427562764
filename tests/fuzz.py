"""Random scenes for fuzz parity (GPU vs oracle, oracle vs numpy model): random settings x adversarial geometry."""
from __future__ import annotations

import numpy as np

from bonnie32_b200 import abi, scenes
from bonnie32_b200.raster import Light, RasterSettings, Texture, Texture15
import cases


def fuzz_scene(seed: int, rgb888: bool = False, n_tris: int = 120):
    rng = np.random.default_rng(seed)
    sc = scenes.scene_c2(n_tris=n_tris, seed=0xF0220000 + seed)
    v = sc.vertices.copy()
    f = sc.faces.copy()
    # geometry: shared vertices, degenerate / huge / behind-camera / non-finite triangles
    kind = rng.integers(0, 8, size=n_tris)
    for t in range(n_tris):
        i = f["v"][t]
        if kind[t] == 0:                                   # re-use vertices of another triangle (indexed mesh)
            f["v"][t] = rng.integers(0, len(v), size=3)
        elif kind[t] == 1:                                 # degenerate: two equal vertices
            v["pos"][i[1]] = v["pos"][i[0]]
        elif kind[t] == 2:                                 # huge triangle reaching far off screen
            v["pos"][i[0]][:2] *= np.float32(rng.choice([50.0, 400.0]))
            v["pos"][i[2]][:2] *= np.float32(-30.0)
        elif kind[t] == 3:                                 # very close to / behind the near plane
            v["pos"][i[rng.integers(0, 3)]][2] = np.float32(rng.choice([0.05, 0.1, 0.1000001, -3.0]))
        elif kind[t] == 4 and rng.random() < 0.03:         # non-finite coordinate: NaN depth (z-mode: draws nothing; sorted lists: panic)
            v["pos"][i[0]][rng.integers(0, 2)] = np.float32(rng.choice([np.inf, -np.inf, np.nan]))
    v["uv"] = (v["uv"] * np.float32(rng.choice([1.0, 3.0])) - np.float32(rng.choice([0.0, 1.5]))).astype(np.float32)
    v["normal"] = rng.normal(size=(len(v), 3)).astype(np.float32)
    v["rgba"][:, 3] = rng.integers(0, 6, size=len(v)) * (rng.random(len(v)) < 0.1)     # Color.blend of vertex colours
    flags_tex = np.where(rng.random(n_tris) < 0.15, abi.FACE_TEX_NONE, rng.integers(0, 4, size=n_tris))    # 3 = out of range
    opaque_only = rng.random() < 0.4
    blend = np.zeros(n_tris, np.int64) if opaque_only else np.where(rng.random(n_tris) < 0.6, 0, rng.integers(1, 6, size=n_tris))
    ea = np.full(n_tris, 255) if opaque_only else np.where(rng.random(n_tris) < 0.8, 255, rng.integers(0, 256, size=n_tris))
    f["flags"] = abi.face_flags(flags_tex, blend, rng.random(n_tris) < 0.5, ea)
    lights = [Light.directional(rng.normal(size=3), float(rng.random() * 1.5)),
              Light.point_colored(rng.normal(size=3) * 10 + np.array([0, 0, 20.0]), float(5 + rng.random() * 60), float(rng.random() * 3),
                                  *rng.random(3))][: rng.integers(0, 3)]
    st = RasterSettings(
        affine_textures=bool(rng.random() < 0.6), use_zbuffer=bool(rng.random() < 0.5), shading=int(rng.integers(0, 3)),
        backface_cull=bool(rng.random() < 0.6), backface_wireframe=False, lights=lights, ambient=float(rng.random() * 1.2),
        dithering=bool(rng.random() < 0.7), wireframe_overlay=False,
        ortho_projection=(float(2 + rng.random() * 10), float(rng.normal()), float(rng.normal())) if rng.random() < 0.15 else None,
        use_rgb555=not rgb888, use_fixed_point=bool(rng.random() < 0.6), xray_mode=bool(rng.random() < 0.15))
    # wireframe phase (its own generator: the scenes of earlier seeds stay what they were).  Not with non-finite
    # coordinates: their saturated end points make edges of billions of steps, which the reference walks for minutes and
    # the device refuses (B32_ERR_UNSUPPORTED; tests/test_gpu_parity.py::test_wireframe_absurd_edge_is_reported_not_walked).
    rng_w = np.random.default_rng(770000 + seed)
    if np.isfinite(v["pos"]).all() and rng_w.random() < 0.3:
        st.backface_wireframe = bool(rng_w.random() < 0.7)
        st.wireframe_overlay = bool(rng_w.random() < 0.4)
    cam = cases._rotated_camera(float(rng.normal() * 0.3), float(rng.normal() * 0.5), rng.normal(size=3) * np.array([2.0, 2.0, 3.0]))
    w, h = [(320, 240), (200, 150), (333, 77), (64, 64), (640, 480)][int(rng.integers(0, 5))]
    fog = (float(rng.random() * 30), float(rng.choice([0.0, 25.0])), float(20 + rng.random() * 60), tuple(int(x) for x in rng.integers(0, 256, size=3))) \
        if (not rgb888 and rng.random() < 0.3) else None
    out = scenes.Scene(f"fuzz{'888' if rgb888 else ''}_{seed}", v, f, [], cam, st, fog=fog, width=w, height=h)
    if rgb888:
        bf = 0.0 if opaque_only else 0.4
        out.textures8 = [cases._rng_texture8(1000 + seed, 32, 32, blend_fraction=bf), cases._rng_texture8(2000 + seed, 8, 64, erase_fraction=0.4, blend_fraction=bf),
                         Texture(0, 0, np.zeros(0, np.uint8))]
    else:
        tb = [0, 0, 0] if opaque_only else [int(x) for x in rng.integers(0, 6, size=3)]
        out.textures = [cases._rng_texture(3000 + seed, 32, 32, blend=tb[0], semi_fraction=0.3), cases._rng_texture(4000 + seed, 8, 64, blend=tb[1], semi_fraction=0.6, zero_fraction=0.3),
                        Texture15(0, 0, np.zeros(0, np.uint16), blend_mode=tb[2])]
    return out
