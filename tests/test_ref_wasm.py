"""The oracle against the reference's own COMPILED code.

tests/golden/ref_wasm/ holds results produced by executing /root/reference/docs/bonnie-32.wasm (the authors' wasm32
release build of the application; `render_mesh_15`, `rasterize_triangle_15`, `project_fixed`, ... are separate named
functions in it) in the interpreter under oracle/wasm/ — see tests/golden/make_ref_wasm.py.  The binary is crate
version 0.1.8 and the source tree 0.1.11; four behavioural differences between them were traced in the decompiled
code (oracle/wasm/DRIFT.md) and are switched back in the oracle by `b32o_set_compat` for this comparison only.
Everything else — transform, snap, near-plane / back-face / fog cull, fog colours, partition + stable sort, edge
stepping, inside test, depth, affine and perspective-correct UV, texel fetch + transparency rules, modulate, Gouraud /
flat multi-light shading, dither, blend modes, x-ray, z-buffer rules, wireframe phase, panics — must agree bit for bit,
framebuffer and z-buffer.
"""
import json
import os

import numpy as np
import pytest

import refbin_cases

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = json.load(open(os.path.join(HERE, "golden", "ref_wasm", "render_mesh_15.json")))
SMALL = {s.name: s for s in refbin_cases.small_scenes()}


@pytest.fixture
def compat_oracle(oracle):
    oracle.lib().b32o_set_compat(refbin_cases.COMPAT_0_1_8)
    yield oracle
    oracle.lib().b32o_set_compat(0)


def _check(oracle, sc):
    rec = FIX["scenes"][sc.name]
    assert rec["inputs"] == refbin_cases.inputs_digest(sc), "scene generator changed: regenerate the fixture"
    rgba, z, tm, rc = oracle.render_scene(sc)
    if "trap" in rec:                       # the reference panicked (NaN sort key / index out of range)
        assert rc != 0, "reference panics, oracle returned OK"
        return
    assert rc == 0
    assert tm["triangles_drawn"] == rec["drawn"]
    a, b = refbin_cases.frame_digest(rgba, z)
    assert a == rec["rgba"], "framebuffer differs from the reference binary"
    assert b == rec["z"], "z-buffer differs from the reference binary"


def test_fixture_covers_every_scene():
    assert set(SMALL) <= set(FIX["scenes"])
    assert len(FIX["wasm_sha256"]) == 64


@pytest.mark.parametrize("name", sorted(SMALL))
def test_oracle_matches_reference_binary(compat_oracle, name):
    _check(compat_oracle, SMALL[name])


@pytest.mark.parametrize("name", [s.name for s in refbin_cases.big_scenes()])
def test_oracle_matches_reference_binary_100k(compat_oracle, name):
    sc = {s.name: s for s in refbin_cases.big_scenes()}[name]
    if sc.name not in FIX["scenes"]:
        pytest.skip("fixture generated without --big")
    _check(compat_oracle, sc)


def test_full_frames_kept_for_debugging(compat_oracle):
    fr = np.load(os.path.join(HERE, "golden", "ref_wasm", "frames.npz"))
    for name in ("c1_single_triangle", "c2_1000_tris_64x64_idx8", "gouraud_lights", "mixed_zbuffer"):
        rgba, z, tm, rc = compat_oracle.render_scene(SMALL[name])
        assert np.array_equal(rgba, fr[name + "/rgba"])
        assert np.array_equal(z.view(np.uint32), fr[name + "/z"].view(np.uint32))


def test_compat_switches_are_off_by_default(oracle):
    assert oracle.lib().b32o_get_compat() == 0
