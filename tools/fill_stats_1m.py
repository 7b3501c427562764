"""Per-tile timeline of k_fill_opaque on the 1M-triangle stress frame (needs build/stats_lib.so, -DB32_FILL_STATS)."""
import sys, os, ctypes as C
sys.path.insert(0, '.')
import __graft_entry__ as g, numpy as np
pkg = g.load_package()
SL = os.path.abspath('build/stats_lib.so'); pkg.abi.LIB_PATH = SL
pkg.abi._lib = None
lib = pkg.abi.load_library(SL)
lib.b32_debug_fill_stats.argtypes = [C.c_void_p, C.c_uint32]
sc = pkg.scenes.scene_c4(n_tris=int(sys.argv[1]) if len(sys.argv) > 1 else 1000000)
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
    print(f"{name:56s} mean {a.mean():8.0f}  p50 {np.median(a):8.0f}  max {a.max():8.0f} ns")
print('kernel span', t2.max() - base, 'ns')
show('candidates done after first CTA start', t0.min(1) - base)
show('first window ready (candidates -> walk order)', (t1 - t0).max(1))
show('  crowd_prepare', (tb0 - t0).max(1))
show('  first crowd_next_window', (tb1 - tb0).max(1))
show('  window sort', (t1 - tb1).max(1))
show('first step landed', (tf - t1).max(1))
show('loop (first data -> last window end)', (tl - tf).max(1))
show('final shade', (t2 - tl).max(1))
show('tile total', t2.max(1) - t0.min(1))
show('batches per warp', nb)
