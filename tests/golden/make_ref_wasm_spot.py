"""Writes tests/golden/ref_wasm/spot.npz and spot.json: Spot lights (render.rs:1038-1059) as the REFERENCE'S OWN COMPILED
code evaluates them (docs/bonnie-32.wasm, build container only):

  acosf                           the binary's `acosf` (compiler_builtins' libm port, func 2058 -> 2057) on 400 000+ arguments:
                                  what `f32::acos` at render.rs:1047 calls in the shipped build
  render::shade_multi_light_color 8 000 (normal, position) pairs x 6 light sets that hold Spot lights
  render_mesh_15 / render_mesh    the scenes of tests/cases.py::spot_scenes / spot_scenes888 (framebuffer + z-buffer digests)

    python tests/golden/make_ref_wasm_spot.py
"""
import hashlib
import json
import os
import struct
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "wasm"))
import __graft_entry__ as g  # noqa: E402

g.load_package()
from bonnie32_b200 import scenes  # noqa: E402
import refbin_cases  # noqa: E402
import refbin_funcs  # noqa: E402
from ref_scene import RefRasterizer  # noqa: E402
from ref_wasm import RefWasm, WasmTrap, WASM  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_wasm")
KEEP_FRAMES = {"spot_flat_mixed_float_nodither"}      # compared with the CUDA output directly (tests/test_gpu_parity.py)
ACOSF = 2058      # two functions carry the name (the export and the libm body it forwards to); the index is unambiguous


def light_record(w, l):
    """Light (60 B): LightType tag@0 + payload@4, name String@36, color@48, intensity@52, enabled@56 (oracle/wasm/ref_scene.py)."""
    name = w.put(b'L')
    payload = [0.0] * 8
    t = int(l.type)
    if t == 0:
        payload[0:3] = [float(x) for x in l.direction]
    elif t == 1:
        payload[0:3] = [float(x) for x in l.position]; payload[3] = float(l.radius)
    else:
        payload[0:3] = [float(x) for x in l.position]; payload[3:6] = [float(x) for x in l.direction]
        payload[6] = float(l.angle); payload[7] = float(l.radius)
    return struct.pack('<I8f', t, *payload) + struct.pack('<III', 1, name, 1) + bytes([0, l.color[0], l.color[1], l.color[2]]) \
        + struct.pack('<f', l.intensity) + bytes([1 if l.enabled else 0, 0, 0, 0])


def main():
    w = RefWasm()
    assert w.func('shade_multi_light_color')
    res = {}
    x = refbin_funcs.acosf_inputs()
    res["acosf"] = np.array([w.call(ACOSF, float(v)) for v in x], np.float32)
    normal, pos, set_idx, ambient = refbin_funcs.spot_shade_inputs()
    set_ptrs = []
    for ls in refbin_funcs.spot_light_sets():
        rec = b''.join(light_record(w, l) for l in ls)
        set_ptrs.append((w.put(rec, 4), len(ls)))
    out = w.alloc(16, 4); npn = w.alloc(12, 4); npp = w.alloc(12, 4)
    shade = np.empty((len(normal), 3), np.float32)
    for i in range(len(normal)):
        w.write(npn, normal[i].tobytes()); w.write(npp, pos[i].tobytes())
        lp, ln = set_ptrs[set_idx[i]]
        w.call('shade_multi_light_color', out, npn, npp, lp, ln, float(ambient[i]))
        shade[i] = np.frombuffer(w.read(out, 12), np.float32)
    res["shade"] = shade
    js = {"wasm_sha256": hashlib.sha256(open(WASM, "rb").read()).hexdigest(), "scenes": {}}
    for sc, rgb888 in [(s, False) for s in refbin_cases.spot_scenes()] + [(s, True) for s in refbin_cases.spot_scenes888()]:
        t = time.time()
        R = RefRasterizer()
        try:
            rgba, z, drawn = R.render_scene888(sc) if rgb888 else R.render_scene(sc, scenes.expand_texture)
            a, b = refbin_cases.frame_digest(rgba, z)
            rec = {"inputs": refbin_cases.inputs_digest(sc), "rgba": a, "z": b, "drawn": int(drawn)}
            if sc.name in KEEP_FRAMES:
                res["frame/" + sc.name] = rgba
        except WasmTrap as e:
            rec = {"inputs": refbin_cases.inputs_digest(sc), "trap": str(e)}
        js["scenes"][sc.name] = rec
        print(f"{sc.name:40s} {time.time() - t:6.2f}s {rec.get('drawn', rec.get('trap'))}", flush=True)
    json.dump(js, open(os.path.join(OUT, "spot.json"), "w"), indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "spot.npz"), **res)
    print("wrote spot.npz", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
