"""Inputs of the overlay-primitive fixtures of the reference binary (tests/golden/make_ref_wasm_prims.py runs them through
docs/bonnie-32.wasm; tests/test_ref_wasm.py replays them through the oracle's b32o_draw_lines).  Deterministic.

Only the primitives that survive as separate functions in the 0.1.8 binary can be called there:
  Framebuffer::draw_line_3d_impl  (render.rs:768-819; allow_equal = draw_line_3d_overlay)
  Framebuffer::draw_circle        (render.rs:631-644)
  Framebuffer::draw_thick_line    (render.rs:875-938; the binary's version takes an unsigned thickness and has no
                                   `thickness <= 1 => draw_line` shortcut, so only thickness >= 2 is comparable)
"""
import numpy as np

from bonnie32_b200 import abi
import cases

# (name, width, height, seed, number of primitives)
CASES = [("prims_320x240", 320, 240, 7101, 400), ("prims_64x48_dense", 64, 48, 7102, 250), ("prims_333x77", 333, 77, 7103, 300),
         ("prims_640x480", 640, 480, 7104, 500)]
KINDS = (abi.LINE_3D, abi.LINE_3D_OVERLAY, abi.LINE_CIRCLE, abi.LINE_THICK)


def primitives(w, h, seed, n):
    rng = np.random.default_rng(seed)
    ln = np.zeros(n, dtype=abi.LINE_DTYPE)
    ln["kind"] = rng.choice(KINDS, n, p=(0.35, 0.35, 0.12, 0.18))
    for f, e in (("x0", w), ("x1", w), ("y0", h), ("y1", h)):
        ln[f] = (e / 2 + (rng.random(n) - 0.5) * e * 1.4).astype(np.int32)           # many leave the screen
    flat = rng.random(n) < 0.15                                                       # horizontals, verticals, points
    ln["x1"] = np.where(flat & (rng.random(n) < 0.5), ln["x0"], ln["x1"])
    ln["y1"] = np.where(flat & (rng.random(n) < 0.5), ln["y0"], ln["y1"])
    planes = np.array([3.0, 4.0, 8.0, 8.0, 12.5, 20.0, 30.0], np.float32)             # on and around line_background's depth planes
    ln["z0"] = rng.choice(planes, n) + (rng.random(n) < 0.3) * rng.normal(0, 2.0, n).astype(np.float32)
    ln["z1"] = np.where(rng.random(n) < 0.4, ln["z0"], rng.choice(planes, n)).astype(np.float32)
    ln["rgb"] = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    ln["blend"] = np.where(rng.random(n) < 0.1, abi.BLEND_ERASE, abi.BLEND_OPAQUE)     # Erase: to_bytes' alpha 0
    circ = ln["kind"] == abi.LINE_CIRCLE
    ln["x1"] = np.where(circ, rng.integers(-2, 14, n), ln["x1"])                      # radius (negative: nothing drawn)
    ln["y1"] = np.where(circ, 0, ln["y1"])
    thick = ln["kind"] == abi.LINE_THICK
    ln["z0"] = np.where(thick, rng.integers(2, 8, n).astype(np.float32), ln["z0"])    # thickness as f32; >= 2: the two versions agree there (oracle/wasm/DRIFT.md)
    ln["z1"] = np.where(thick | circ, np.float32(0), ln["z1"])
    ln["z0"] = np.where(circ, np.float32(0), ln["z0"])
    return ln


def background(w, h, seed):
    return cases.line_background(w, h, seed)
