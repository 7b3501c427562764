"""The conservative trivial reject of surfaces whose edge values are the reference's rounded incremental additions
(b32_kernels.cu: stepped_surface_misses_box; render.rs:1517-1542, 1706-1712): a numpy mirror of the device arithmetic
(float32, one rounding per operator, the same operation order) against brute-force chains.  A box may only be rejected when
every one of its pixels fails the reference's inside test — over random float-projected surfaces, thin and huge triangles,
far off-screen vertices, and boxes of the two block shapes the fill uses (8x4, 4x4) plus whole tiles."""
import numpy as np
import pytest

F = np.float32
ERR = F(-0.0001)


def chain_values(w_start, a, b, nx, ny):
    """w[dy, dx] as the reference steps it: row starts by repeated + b, then along the row by repeated + a."""
    out = np.empty((ny, nx), F)
    row = F(w_start)
    for dy in range(ny):
        w = row
        for dx in range(nx):
            out[dy, dx] = w
            w = F(w + a)
        row = F(row + b)
    return out


def mirror_reject(s, ix0, ix1, iy0, iy1):
    """stepped_surface_misses_box, operator for operator (ix*, iy* are offsets from the bbox origin, inclusive)."""
    w0s, w1s, a0, b0, a1, b1, inv = (F(s[k]) for k in ("w0s", "w1s", "a0", "b0", "a1", "b1", "inv_area"))
    dx0, dx1, dy0, dy1 = F(ix0), F(ix1), F(iy0), F(iy1)
    with np.errstate(all="ignore"):
        r0 = F(w0s + F(dy0 * b0)); r1 = F(w0s + F(dy1 * b0)); q0 = F(w1s + F(dy0 * b1)); q1 = F(w1s + F(dy1 * b1))
        ax0 = F(dx0 * a0); ax1 = F(dx1 * a0); cx0 = F(dx0 * a1); cx1 = F(dx1 * a1)
        xs = [F(F(r0 + ax0) * inv), F(F(r0 + ax1) * inv), F(F(r1 + ax0) * inv), F(F(r1 + ax1) * inv)]
        ys = [F(F(q0 + cx0) * inv), F(F(q0 + cx1) * inv), F(F(q1 + cx0) * inv), F(F(q1 + cx1) * inv)]
        xmax, xmin, ymax, ymin = max(xs), min(xs), max(ys), min(ys)
        steps = F(F(ix1 + iy1 + 8) * F(2.384185791015625e-07))
        ia = abs(inv)
        ex = F(F(steps * F(F(abs(w0s) + F(dy1 * abs(b0))) + F(dx1 * abs(a0)))) * ia)
        ey = F(F(steps * F(F(abs(w1s) + F(dy1 * abs(b1))) + F(dx1 * abs(a1)))) * ia)
        slack = F(F(9.5367431640625e-07) * F(F(F(F(F(1.0) + abs(xmax)) + abs(xmin)) + abs(ymax)) + abs(ymin)))
        cz = F(F(F(F(F(1.0) - xmin) - ymin) + ex) + ey) + F(F(3.0) * slack)
        return bool(F(F(xmax + ex) + slack) < ERR or F(F(ymax + ey) + slack) < ERR or F(cz) < ERR)


def surface_from_triangle(v1, v2, v3, width, height):
    """Edge setup of rasterize_triangle_15 (render.rs:1455-1518) for screen-space vertices (x, y)."""
    (x1, y1), (x2, y2), (x3, y3) = [(F(p[0]), F(p[1])) for p in (v1, v2, v3)]
    min_x = int(max(min(x1, x2, x3), F(0))); max_x = int(min(F(max(x1, x2, x3) + F(1)), F(width)))
    min_y = int(max(min(y1, y2, y3), F(0))); max_y = int(min(F(max(y1, y2, y3) + F(1)), F(height)))
    if min_x >= max_x or min_y >= max_y:
        return None
    area = F(F(F(y2 - y3) * F(x1 - x3)) + F(F(x3 - x2) * F(y1 - y3)))
    if abs(area) < F(1e-5):
        return None
    s = dict(inv_area=F(F(1.0) / area), a0=F(y2 - y3), b0=F(x3 - x2), a1=F(y3 - y1), b1=F(x1 - x3), min_x=min_x, max_x=max_x, min_y=min_y, max_y=max_y)
    sx, sy = F(min_x), F(min_y)
    s["w0s"] = F(F(s["a0"] * F(sx - x3)) + F(s["b0"] * F(sy - y3)))
    s["w1s"] = F(F(s["a1"] * F(sx - x3)) + F(s["b1"] * F(sy - y3)))
    return s


def inside_mask(s):
    nx, ny = s["max_x"] - s["min_x"], s["max_y"] - s["min_y"]
    w0 = chain_values(s["w0s"], s["a0"], s["b0"], nx, ny); w1 = chain_values(s["w1s"], s["a1"], s["b1"], nx, ny)
    with np.errstate(all="ignore"):
        bx = (w0 * s["inv_area"]).astype(F); by = (w1 * s["inv_area"]).astype(F)
        bz = ((F(1.0) - bx).astype(F) - by).astype(F)
    return (bx >= ERR) & (by >= ERR) & (bz >= ERR)


def random_triangle(rng, width, height):
    kind = rng.integers(0, 5)
    if kind == 0:        # level-sized
        c = rng.random(2) * [width, height]; return [c + rng.normal(size=2) * 40 for _ in range(3)]
    if kind == 1:        # thin sliver
        c = rng.random(2) * [width, height]; d = rng.normal(size=2) * 120
        return [c, c + d, c + d * rng.random() + rng.normal(size=2) * 0.7]
    if kind == 2:        # far off-screen vertices
        return [rng.normal(size=2) * 3000 + [width / 2, height / 2] for _ in range(3)]
    if kind == 3:        # huge, with one enormous coordinate
        return [rng.normal(size=2) * 60000, rng.random(2) * [width, height], rng.normal(size=2) * 500]
    return [rng.random(2) * [width, height] for _ in range(3)]     # screen-sized


@pytest.mark.parametrize("seed", range(6))
def test_conservative_reject_never_drops_an_inside_pixel(seed):
    rng = np.random.default_rng(4400 + seed)
    width, height = [(320, 240), (333, 77), (640, 480)][seed % 3]
    n_rejected = n_boxes = n_empty_boxes = 0
    for _ in range(60):
        s = surface_from_triangle(*random_triangle(rng, width, height), width, height)
        if s is None:
            continue
        inside = inside_mask(s)
        for bw, bh in ((8, 4), (4, 4), (16, 16)):
            for _ in range(40):
                # a block of the screen grid that overlaps the bbox, clipped to it
                gx = rng.integers(s["min_x"] // bw, (s["max_x"] - 1) // bw + 1) * bw
                gy = rng.integers(s["min_y"] // bh, (s["max_y"] - 1) // bh + 1) * bh
                ix0 = max(gx, s["min_x"]) - s["min_x"]; ix1 = min(gx + bw, s["max_x"]) - 1 - s["min_x"]
                iy0 = max(gy, s["min_y"]) - s["min_y"]; iy1 = min(gy + bh, s["max_y"]) - 1 - s["min_y"]
                any_inside = bool(inside[iy0:iy1 + 1, ix0:ix1 + 1].any())
                rej = mirror_reject(s, ix0, ix1, iy0, iy1)
                n_boxes += 1; n_rejected += rej; n_empty_boxes += not any_inside
                assert not (rej and any_inside), (s, ix0, ix1, iy0, iy1)
    assert n_rejected > 0.5 * n_empty_boxes > 0            # and it is worth having: most empty boxes are rejected
