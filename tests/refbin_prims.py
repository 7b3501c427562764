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


# ---- plain draw_line (render.rs:715-751) through draw::draw_3d_line_clipped (draw.rs:12-66), the one caller that survives
# as a separate function: world-space segments in front of an identity camera at the origin.
CLIPPED = [("lines_320x240", 320, 240, 7201, 300), ("lines_97x61", 97, 61, 7202, 200)]


def clipped_segments(w, h, seed, n):
    """(p0[n,3], p1[n,3], rgb[n,3]) in world space, and the integer end points world_to_screen (math.rs:503-534) gives them
    with the default camera at the origin — Camera::new's basis is x = (-1, 0, 0), y = (0, -1, 0), z = (0, 0, 1)
    (camera.rs:76-91), so cam_x = -x, cam_y = -y, cam_z = z exactly — every operation in f32, in the reference's order."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    z = (1.0 + 99.0 * rng.random((n, 2))).astype(f32)
    span = 1.6                                             # end points up to 60 % outside the screen
    vs = f32(f32(min(w, h)) / f32(2.0)) * f32(0.75)
    xy = ((rng.random((n, 2, 2)) - 0.5) * span).astype(f32)
    p = np.zeros((n, 2, 3), f32)
    for k in range(2):
        denom = z[:, k] + f32(5.0)
        p[:, k, 0] = xy[:, k, 0] * f32(w) * denom / (f32(4.0) * vs)      # roughly on-screen x = (xy + 0.5) * w
        p[:, k, 1] = xy[:, k, 1] * f32(h) * denom / (f32(4.0) * vs)
        p[:, k, 2] = z[:, k]
    ends = np.zeros((n, 2, 2), np.int32)
    for k in range(2):
        denom = p[:, k, 2] + f32(5.0)
        sx = (-p[:, k, 0] * f32(4.0) / denom) * vs + f32(f32(w) / f32(2.0))
        sy = (-p[:, k, 1] * f32(4.0) / denom) * vs + f32(f32(h) / f32(2.0))
        ends[:, k, 0] = np.trunc(sx).astype(np.int32)                    # `as i32`
        ends[:, k, 1] = np.trunc(sy).astype(np.int32)
    rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    return p[:, 0].copy(), p[:, 1].copy(), rgb, ends
