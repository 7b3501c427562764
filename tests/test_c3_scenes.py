"""BASELINE config 3: the reference's sample world scenes through the full per-room pipeline
(scene.rs render_scene: one render_mesh_15 per room, per-room ambient + fog, Gouraud, RGB555 +
dither), z-buffer ON (shipped default) and painter's mode.  Fixtures: tests/golden/make_c3.py."""
import hashlib
import json
import os

import numpy as np
import pytest

import bonnie32_b200 as pkg
import c3

HASHES = json.load(open(os.path.join(c3.GOLDEN, "c3_hashes.json")))
PATHS = c3.scene_paths()
CASES = [(p, m) for p in PATHS for m in c3.MODES]
IDS = [f"{os.path.basename(p)[3:-4]}-{m}" for p, m in CASES]


def render_oracle(oracle, sc, kw):
    rgba = np.empty((sc.height, sc.width, 4), np.uint8)
    z = np.empty((sc.height, sc.width), np.float32)
    rgba[...] = np.array(list(sc.clear) + [255], np.uint8)
    z[...] = np.finfo(np.float32).max
    drawn = 0
    for rc in sc.rooms:
        rcode, tm, _ = oracle.render_mesh_15(rgba, z, rc.vertices, rc.faces, sc.textures, sc.camera, sc.settings(rc.ambient, **kw), rc.fog)
        assert rcode == 0
        drawn += tm["triangles_drawn"]
    from bonnie32_b200.raster import libm_cosf, libm_sinf
    for pc in sc.parts:                              # placed asset parts, scene.rs:109-169 (the oracle's own restatement of the transform)
        moved = abs(pc.facing) > 0.0001 or any(abs(x) > 0.0001 for x in pc.world_pos)
        v = oracle.place_vertices(pc.vertices, pc.facing, libm_cosf(pc.facing), libm_sinf(pc.facing), pc.world_pos) if moved else pc.vertices
        rcode, tm, _ = oracle.render_mesh_15(rgba, z, v, pc.faces, sc.textures, sc.camera, sc.part_settings(pc, **kw), pc.fog)
        assert rcode == 0
        drawn += tm["triangles_drawn"]
    return rgba, z, drawn


def test_fixtures_present():
    assert len(PATHS) == 6
    sc = c3.load_scene(PATHS[0])
    assert sc.rooms and sc.textures and len(sc.rooms[0].vertices) > 0


@pytest.mark.parametrize("path,mode", CASES, ids=IDS)
def test_oracle_matches_numpy_model_golden(oracle, path, mode):
    sc = c3.load_scene(path)
    rgba, z, drawn = render_oracle(oracle, sc, c3.MODES[mode])
    want = HASHES[f"{sc.name}:{mode}"]
    assert drawn == want["triangles_drawn"]
    assert hashlib.sha256(rgba.tobytes()).hexdigest() == want["rgba_sha256"]
    assert hashlib.sha256(z.tobytes()).hexdigest() == want["z_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("path,mode", CASES, ids=IDS)
def test_gpu_matches_oracle(ctx, oracle, path, mode):
    sc = c3.load_scene(path)
    want, want_z, drawn = render_oracle(oracle, sc, c3.MODES[mode])
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    fb.clear(sc.clear)
    got_drawn = 0
    for rc in sc.rooms:                      # several calls compose on one device framebuffer
        tm = pkg.render_mesh_15(fb, rc.vertices, rc.faces, sc.textures, sc.camera, sc.settings(rc.ambient, **c3.MODES[mode]), rc.fog)
        got_drawn += tm["triangles_drawn"]
    for pc in sc.parts:                      # placed asset parts: resident part + per-object transform on the device (b32_render_mesh_placed)
        mesh = pkg.Mesh(ctx, pc.vertices, pc.faces)
        tm = mesh.render_placed(sc.camera, sc.part_settings(pc, **c3.MODES[mode]), pc.facing, pc.world_pos, pc.fog)
        got_drawn += tm["triangles_drawn"]
        mesh.free()
    got, got_z = fb.download()
    assert got_drawn == drawn
    bad = (got != want).any(-1)
    assert not bad.any(), f"{sc.name}/{mode}: {bad.sum()} pixels differ"
    assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32))
    assert hashlib.sha256(got.tobytes()).hexdigest() == HASHES[f"{sc.name}:{mode}"]["rgba_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", PATHS, ids=[os.path.basename(p)[3:-4] for p in PATHS])
def test_gpu_level_renderer_cached_geometry(ctx, oracle, path):
    """levels.LevelRenderer: rooms uploaded once per level generation, every frame enqueued (frame graphs), several
    frames with a moving camera — each equals the oracle's per-room loop."""
    from bonnie32_b200 import levels
    import cases
    sc = c3.load_scene(path)
    fb = pkg.Framebuffer(sc.width, sc.height, ctx)
    lr = levels.LevelRenderer(ctx, sc)
    for k in range(4):
        cam = cases._rotated_camera(0.0, 0.0, (0.0, 0.0, 0.0))
        cam.position = (sc.camera.position + np.float32(k * 37.0) * sc.camera.basis_x).astype(np.float32)
        cam.basis_x, cam.basis_y, cam.basis_z = sc.camera.basis_x, sc.camera.basis_y, sc.camera.basis_z
        lr.render(fb, cam)
        got, got_z = fb.download()
        moved = type(sc)(sc.name, sc.rooms, sc.textures, cam, parts=sc.parts, lights=sc.lights)
        want, want_z, drawn = render_oracle(oracle, moved, {})
        assert np.array_equal(got, want), (sc.name, k)
        assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32)), (sc.name, k)
    lr.generation += 1                      # an edit: geometry is uploaded again
    lr.render(fb)
    got, _ = fb.download()
    want, _, _ = render_oracle(oracle, sc, {})
    assert np.array_equal(got, want)
    lr.close()


def test_scene_lights_and_placed_parts_assembly(oracle):
    """collect_scene_lights (scene.rs:32-70) and the object loop (scene.rs:219-259) on a hand-built level: per-instance
    overrides, disabled instances, a missing asset, world_position with and without a floor, fan triangulation."""
    from bonnie32_b200 import levels, abi
    F = np.float32
    room = {"position": {"x": F(-2048.0), "y": F(512.0), "z": F(1024.0)}, "ambient": F(0.25),
            "sectors": [[None, {"floor": {"heights": (F(0.0), F(256.0), F(256.0), F(0.0))}}], [None, None]],
            "fog": {"enabled": False},
            "objects": [{"sector_x": 0, "sector_z": 1, "height": F(100.0), "facing": F(0.5), "asset_id": 7, "enabled": True},
                        {"sector_x": 1, "sector_z": 0, "height": F(0.0), "facing": F(0.0), "asset_id": 7, "enabled": True,
                         "overrides": {"light": {"color": (10, 20, 30), "intensity": F(3.0), "radius": None, "offset": None}}},
                        {"sector_x": 0, "sector_z": 0, "height": F(0.0), "facing": F(0.0), "asset_id": 7, "enabled": False},
                        {"sector_x": 0, "sector_z": 0, "height": F(0.0), "facing": F(0.0), "asset_id": 99, "enabled": True}]}
    quad = {"vertices": [{"pos": {"x": F(x), "y": F(y), "z": F(0.0)}, "uv": {"x": F(u), "y": F(v)}, "normal": {"x": F(0), "y": F(0), "z": F(1)},
                          "color": {"r": 200, "g": 100, "b": 50, "blend": "Opaque"}} for x, y, u, v in ((0, 0, 0, 1), (64, 0, 1, 1), (64, 64, 1, 0), (0, 64, 0, 0), (32, 96, 0.5, 0))],
            "faces": [{"vertices": [0, 1, 2, 3, 4], "texture_id": None, "black_transparent": False, "blend_mode": "Add"},
                      {"vertices": [0, 1], "texture_id": None, "black_transparent": True, "blend_mode": "Opaque"}]}
    asset = {"id": 7, "components": [{"__variant__": "Light", "value": {"color": (255, 128, 0), "intensity": F(1.5), "radius": F(4096.0), "offset": (F(0.0), F(512.0), F(0.0))}},
                                     {"__variant__": "Mesh", "value": {"parts": [{"name": "p", "mesh": quad, "texture_ref": "Checkerboard", "visible": True, "double_sided": True},
                                                                                   {"name": "hidden", "mesh": quad, "visible": False}]}}]}
    level = {"rooms": [room]}
    lights = levels.collect_scene_lights(level, {7: asset})
    assert len(lights) == 2
    # instance 0: floor average (0 + 256 + 256 + 0) / 4 = 128 (not offset by room.y, as in the reference) + height 100, light offset +512 in y
    assert np.allclose(lights[0].position, (-2048.0 + 512.0, 128.0 + 100.0 + 512.0, 1024.0 + 1024.0 + 512.0))
    assert lights[0].color == (255, 128, 0) and lights[0].intensity == 1.5 and lights[0].radius == 4096.0
    # instance 1: no floor -> room.position.y; colour and intensity overridden, radius and offset from the asset
    assert np.allclose(lights[1].position, (-2048.0 + 1024.0 + 512.0, 512.0 + 512.0, 1024.0 + 512.0))
    assert lights[1].color == (10, 20, 30) and lights[1].intensity == 3.0 and lights[1].radius == 4096.0
    parts, texs = levels.assemble_parts(level, {7: asset}, {}, first_tex=5)
    assert len(parts) == 2 and len(texs) == 2                      # two enabled instances with a known asset x one visible part
    pc = parts[0]
    assert len(pc.vertices) == 5 and len(pc.faces) == 3            # pentagon -> fan of 3; the 2-vertex face is dropped
    assert pc.faces["v"].tolist() == [[0, 1, 2], [0, 2, 3], [0, 3, 4]]
    assert (pc.faces["flags"] & 0xFFFF).tolist() == [5, 5, 5] and ((pc.faces["flags"] >> 16) & 7).tolist() == [abi.BLEND_ADD] * 3
    assert ((pc.faces["flags"] >> 19) & 1).tolist() == [0, 0, 0] and pc.double_sided and pc.facing == 0.5
    assert texs[0].width == 0                                       # no atlas in the file: an empty texture (sample() = transparent)
    assert (parts[1].faces["flags"] & 0xFFFF).tolist() == [6, 6, 6]
