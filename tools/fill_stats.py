"""Per-warp timeline of k_fill_opaque on the C4 frame (GPU box only; needs a -DB32_FILL_STATS build):
    tools/build_variant.sh build/stats_lib.so -DB32_FILL_STATS && python tools/fill_stats.py"""
import sys, os, ctypes as C
sys.path.insert(0, '.')
import __graft_entry__ as g, numpy as np
pkg = g.load_package()
SL = os.path.abspath('build/stats_lib.so'); pkg.abi.LIB_PATH = SL
pkg.abi._lib = None
lib = pkg.abi.load_library(SL)
lib.b32_debug_fill_stats.argtypes = [C.c_void_p, C.c_uint32]
sc = pkg.scenes.scene_c4()
ctx = pkg.Context(0)
fb = pkg.Framebuffer(320, 240, ctx)
ctx.set_textures(sc.textures)
mesh = pkg.Mesh(ctx, sc.vertices, sc.faces)
for _ in range(3):
    fb.clear(sc.clear); mesh.render(sc.camera, sc.settings)
st = np.zeros(300 * 16 * 8, np.uint32)
lib.b32_debug_fill_stats(st.ctypes.data, st.size)
st = st.reshape(300, 16, 8).astype(np.int64)
t0, t1, t2, nb, tf, tb0, tb1, tl = (st[:, :, k] for k in range(8))
base = t0.min()
def show(name, a):
    print(f"{name:46s} mean {a.mean():8.0f}  p50 {np.median(a):8.0f}  p90 {np.percentile(a, 90):8.0f}  max {a.max():8.0f} ns")
print('kernel span (first CTA start -> last warp end)', t2.max() - base, 'ns')
show('CTA start after first CTA start', t0.min(1) - base)
show('sort (start -> walk order ready), per tile', (t1 - t0).max(1))
show('first step landed (sorted -> first barrier)', (tf - t1).max(1))
show('batch 0 (first data -> end of batch 0)', tb0 - tf)
two = tb1 > 0
show('batch 1 (warps that walk >= 3 batches)', (tb1 - tb0)[two])
show('loop (first data -> loop end, incl. waiting for the CTA)', tl - tf)
show('final shade + store', t2 - tl)
show('tile total (start -> last warp end)', t2.max(1) - t0.min(1))
show('32-entry batches walked per warp', nb)
print('batches histogram', np.bincount(nb.reshape(-1)))
# slowest tiles: where does their time go?
tot = t2.max(1) - t0.min(1)
order = np.argsort(-tot)[:8]
print('slowest tiles: tile, total ns, sort ns, max batches, mean batches, loop ns (max warp), shade ns (max warp), start offset')
for t in order:
    print(int(t), int(tot[t]), int((t1 - t0)[t].max()), int(nb[t].max()), round(float(nb[t].mean()), 1), int((tl - tf)[t].max()), int((t2 - tl)[t].max()), int(t0[t].min() - base))
print('corr(total, max batches) =', round(float(np.corrcoef(tot, nb.max(1))[0, 1]), 3), ' corr(total, start offset) =', round(float(np.corrcoef(tot, t0.min(1) - base)[0, 1]), 3))
