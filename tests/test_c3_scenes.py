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
        moved = type(sc)(sc.name, sc.rooms, sc.textures, cam)
        want, want_z, drawn = render_oracle(oracle, moved, {})
        assert np.array_equal(got, want), (sc.name, k)
        assert np.array_equal(got_z.view(np.uint32), want_z.view(np.uint32)), (sc.name, k)
    lr.generation += 1                      # an edit: geometry is uploaded again
    lr.render(fb)
    got, _ = fb.download()
    want, _, _ = render_oracle(oracle, sc, {})
    assert np.array_equal(got, want)
    lr.close()
