"""Writes tests/golden/ref_wasm/levels.json: sha256 of every sample room's triangles as produced by the REFERENCE BINARY
(docs/bonnie-32.wasm: load_level_from_str + Room::add_*_to_render_data, see oracle/wasm/ref_level.py), and checks on the spot
that bonnie-32_b200/levels.py produces the same vertices (position, uv, normal, colour + blend tag) and faces (indices,
black_transparent) bit for bit.  Build container only.

    python tests/golden/make_ref_wasm_levels.py
"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "wasm"))
import __graft_entry__ as g  # noqa: E402

g.load_package()
from bonnie32_b200 import levels  # noqa: E402
import ref_level  # noqa: E402
import refbin_cases  # noqa: E402

LEVELS = "/root/reference/assets/samples/levels"


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(LEVELS, "*.ron"))):
        name = os.path.splitext(os.path.basename(path))[0]
        ref = ref_level.level_geometry(levels.brotli_decompress(open(path, "rb").read()))
        lv = levels.load_level_file(path)
        assert len(ref) == len(lv["rooms"])
        rooms = []
        for room, (rv, rf) in zip(lv["rooms"], ref):
            n = len(rv["blend"])
            rgba = np.concatenate([rv["rgb"], rv["blend"][:, None]], axis=1)
            d_ref = refbin_cases.geometry_digest(rv["pos"].reshape(n, 3), rv["uv"].reshape(n, 2), rv["normal"].reshape(n, 3), rgba, rf["v"], rf["black_transparent"])
            v, f = levels.room_to_render_data(room, lambda r: None)
            d_own = refbin_cases.geometry_digest(v["pos"], v["uv"], v["normal"], v["rgba"], f["v"], ((f["flags"] >> 19) & 1).astype(np.uint8))
            assert d_ref == d_own, f"{name}: levels.py differs from the reference binary"
            d_nouv = refbin_cases.geometry_digest(rv["pos"].reshape(n, 3), None, rv["normal"].reshape(n, 3), rgba, rf["v"], rf["black_transparent"], with_uv=False)
            rooms.append({"vertices": int(n), "faces": int(len(rf["v"])), "sha256": d_ref, "sha256_no_uv": d_nouv})
            print(name, n, len(rf["v"]), d_ref[:16])
        out[name] = rooms
    json.dump({"levels": out, "how": "docs/bonnie-32.wasm: load_level_from_str + Room::add_horizontal_face/_wall/_diagonal_wall_to_render_data, "
                                     "resolver = miss (texture 0, width 64); sha256 over pos, uv, normal, rgba, face indices, black_transparent"},
              open(os.path.join(ROOT, "tests", "golden", "ref_wasm", "levels.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
