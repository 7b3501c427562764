"""Inputs of the direct function-level fixtures of the reference binary (tests/golden/make_ref_wasm_funcs.py writes the
outputs; tests/test_ref_wasm.py replays the inputs through the oracle).  Deterministic (SplitMix64 / fixed lists)."""
import numpy as np

from bonnie32_b200 import scenes
from bonnie32_b200.raster import Light
import cases

SIZES = [(320, 240), (640, 480), (333, 77)]


def cameras():
    return [cases._rotated_camera(rx, ry, pos) for rx, ry, pos in
            [(0.0, 0.0, (0, 0, 0)), (0.3, 0.7, (10.0, -20.0, 300.0)), (-0.2, 2.4, (-30.0, 15.0, 900.0)), (1.2, -3.0, (0.5, 0.25, -4.0)),
             (0.01, 0.02, (1e-3, -1e-3, 1e-4)), (-1.5, 0.0, (1000.0, 2000.0, -3000.0)), (0.7, 5.5, (-7.0, 3.0, 2.0)), (3.1, 3.1, (100000.0, 0.0, 0.0))]]


def project_inputs(n=24000):
    u = scenes.splitmix64_u01(0xF1ED0001, n * 5).reshape(n, 5)
    scale = np.choose((u[:, 3] * 6).astype(int), [1.0, 10.0, 100.0, 3000.0, 0.01, 600000.0])
    world = ((2.0 * u[:, :3] - 1.0) * scale[:, None]).astype(np.float32)
    cam_idx = (u[:, 4] * 8).astype(np.int32) % 8
    size_idx = (u[:, 3] * 977).astype(np.int32) % 3
    # adversarial rows: non-finite, denormal, exactly on / next to the |denom| < 256 early-out (camera 0: denom = z + 5)
    special = np.array([[np.nan, 0, 1], [0, np.inf, 1], [1, 2, -np.inf], [1e-42, -1e-42, 1e-40], [3.4e38, -3.4e38, 3.4e38],
                        [1, 1, -5.0], [1, 1, -5.0 + 255.0 / 4096.0], [1, 1, -5.0 + 256.0 / 4096.0], [1, 1, -5.0 - 255.0 / 4096.0],
                        [1, 1, -5.0 - 256.0 / 4096.0], [524287.9, -524288.0, 524287.0], [524288.0, 524288.5, -524289.0],
                        [0.5 / 4096, 1.5 / 4096, 2.5 / 4096], [-0.5 / 4096, -1.5 / 4096, -2.5 / 4096]], np.float32)
    world[:len(special)] = special
    cam_idx[:len(special)] = 0
    return world, cam_idx, size_idx


def light_sets():
    off = Light.point((0.0, 0.0, 10.0), 50.0, 2.0)
    off.enabled = False
    return [
        [],
        [Light.directional((-1.0, -1.0, -1.0), 0.7)],
        [Light.directional((0.3, -2.0, 0.5), 1.4), Light.point_colored((0.5, 0.5, 4.0), 30.0, 1.5, 1.0, 0.5, 0.25), Light.point((-3.0, 2.0, 20.0), 25.0, 0.9), off],
        [Light.point((0.0, 0.0, 0.0), 0.0, 1.0), Light.point((1.0, 2.0, 3.0), 1e-3, 5.0), Light.point_colored((5.0, 5.0, 5.0), 1e6, 100.0, 0.1, 0.9, 0.5)],
        [Light.directional((0.0, 0.0, 0.0), 1.0), Light.directional((1e-20, 0.0, 0.0), 2.0)],
        [Light.point((1.0, 2.0, 3.0), 10.0, -1.0), Light.directional((0.0, 1.0, 0.0), -0.5)],
    ]


def shade_inputs(n=6000):
    u = scenes.splitmix64_u01(0xF1ED0002, n * 8).reshape(n, 8)
    normal = (2.0 * u[:, :3] - 1.0).astype(np.float32)
    pos = ((2.0 * u[:, 3:6] - 1.0) * 30.0).astype(np.float32)
    set_idx = (u[:, 6] * 6).astype(np.int32) % 6
    ambient = (u[:, 7] * 1.2).astype(np.float32)
    # a few exact hits: position == light position (dist < 0.001), zero normal, NaN normal
    pos[0] = (1.0, 2.0, 3.0); set_idx[0] = 3
    pos[1] = (0.5, 0.5, 4.0); set_idx[1] = 2
    normal[2] = 0.0
    normal[3] = (np.nan, 0.0, 1.0)
    pos[4] = (np.inf, 0.0, 0.0)
    return normal, pos, set_idx, ambient


def blend_inputs(n=30000):
    """(color15 incl. bit 15, back r/g/b, blend mode) for Framebuffer::set_pixel_blended_15."""
    u = scenes.splitmix64_u01(0xF1ED0003, n * 5).reshape(n, 5)
    c15 = np.floor(u[:, 0] * 65536.0).astype(np.uint16)
    back = np.floor(u[:, 1:4] * 256.0).astype(np.uint8)
    mode = (np.floor(u[:, 4] * 6.0).astype(np.uint8)) % 6
    # a few corners: saturating add, clamping subtract, black front / back
    c15[:6] = [0xFFFF, 0x8000, 0x8001, 0xFC00, 0x83FF, 0x0000]
    back[:6] = [[255, 255, 255], [0, 0, 0], [7, 8, 248], [255, 0, 128], [16, 31, 249], [1, 2, 3]]
    return c15, back, mode


def acosf_inputs(n=400000):
    """Arguments for the binary's `acosf` (compiler_builtins libm): uniform over [-1.0001, 1.0001], denser towards 0 and
    towards +-1 / +-0.5 (the branch boundaries), every float within 64 ulps of the boundaries, and the specials."""
    u = scenes.splitmix64_u01(0xF1ED0004, n * 2).reshape(n, 2)
    x = (2.0002 * u[:, 0] - 1.0001)
    k = (u[:, 1] * 5).astype(int)
    x = np.where(k == 1, x ** 5, x)                                    # near 0 (down to the 2^-26 early-out)
    x = np.where(k == 2, np.sign(x) * (1.0 - np.abs(x) ** 6 * 1e-3), x)    # just inside +-1
    x = np.where(k == 3, np.sign(x) * (0.5 + x ** 7 * 1e-3), x)        # around +-0.5
    x = x.astype(np.float32)
    near = []
    for b in (0.0, 0.5, 1.0, 2.0 ** -26, 2.0 ** -27, 1e-38, 0.70710678):
        bits = int(np.float32(b).view(np.uint32))
        w = np.arange(max(bits - 64, 0), bits + 65, dtype=np.uint32)
        near += [w, w | np.uint32(0x80000000)]
    special = np.array([np.nan, np.inf, -np.inf, 2.0, -2.0, 3.4e38, -3.4e38, 1e-45, -1e-45, 1.0000001, -1.0000001], np.float32)
    return np.concatenate([x, np.concatenate(near).view(np.float32), special])


def spot_light_sets():
    import cases
    sp = cases.spot_lights()
    neg_angle = Light.spot((1.0, 2.0, 3.0), (0.0, 1.0, 0.0), -0.5, 40.0, 1.0)
    wide = Light.spot((0.0, 0.0, 0.0), (1.0, 0.0, 0.0), 3.2, 100.0, 1.0)              # cone wider than pi: everything inside the radius
    zero_dir = Light.spot((2.0, 2.0, 2.0), (0.0, 0.0, 0.0), 1.0, 60.0, 2.0)           # normalize(0) = 0: acos(0) = pi/2 > 1.0
    zero_dir2 = Light.spot((2.0, 2.0, 2.0), (0.0, 0.0, 0.0), 1.6, 60.0, 2.0)
    tiny_r = Light.spot((1.0, 2.0, 3.0), (0.0, 0.0, 1.0), 1.0, 1e-3, 5.0)
    return [
        [sp[0]],
        sp,
        [Light.directional((0.3, -2.0, 0.5), 1.4), sp[1], Light.point((-3.0, 2.0, 20.0), 25.0, 0.9), sp[2], sp[4]],
        [neg_angle, wide, zero_dir, zero_dir2],
        [tiny_r, sp[3], wide],
        [sp[2]],
    ]


def spot_shade_inputs(n=8000):
    u = scenes.splitmix64_u01(0xF1ED0005, n * 8).reshape(n, 8)
    normal = (2.0 * u[:, :3] - 1.0).astype(np.float32)
    pos = ((2.0 * u[:, 3:6] - 1.0) * 40.0).astype(np.float32)
    set_idx = (u[:, 6] * 6).astype(np.int32) % 6
    ambient = (u[:, 7] * 1.2).astype(np.float32)
    # on the axis of the torch (dot = -1 exactly, and a hair beyond with the un-normalised light), at a light's position
    pos[0] = (0.0, 0.0, 10.0); set_idx[0] = 0
    pos[1] = (0.0, 0.0, -10.0); set_idx[1] = 0
    pos[2] = (1.0, 2.0, 3.0); set_idx[2] = 3
    pos[3] = (5.0 - 0.2 * 7, -20.0 + 1.1 * 7, 10.0 + 1.3 * 7); set_idx[3] = 5
    normal[4] = 0.0
    normal[5] = (np.nan, 0.0, 1.0)
    pos[6] = (np.inf, 0.0, 0.0)
    return normal, pos, set_idx, ambient
