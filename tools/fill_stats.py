"""Per-warp timeline of k_fill_opaque on the C4 frame (GPU box only; needs a -DB32_FILL_STATS build):
    tools/build_variant.sh stats_lib.so -DB32_FILL_STATS && python tools/fill_stats.py"""
import sys, ctypes as C
sys.path.insert(0,'.')
import __graft_entry__ as g, numpy as np
pkg=g.load_package()
import os; SL=os.path.abspath('stats_lib.so'); pkg.abi.LIB_PATH=SL
pkg.abi._lib=None
lib=pkg.abi.load_library(SL)
lib.b32_debug_fill_stats.argtypes=[C.c_void_p, C.c_uint32]
sc=pkg.scenes.scene_c4()
ctx=pkg.Context(0)
fb=pkg.Framebuffer(320,240,ctx)
ctx.set_textures(sc.textures)
mesh=pkg.Mesh(ctx, sc.vertices, sc.faces)
for _ in range(3):
    fb.clear(sc.clear); mesh.render(sc.camera, sc.settings)
st=np.zeros(300*16*8, np.uint32)
lib.b32_debug_fill_stats(st.ctypes.data, st.size)
st=st.reshape(300,16,8).astype(np.int64)
t0=st[:,:,0]; t1=st[:,:,1]; t2=st[:,:,2]
base=t0.min()
sort_ns=(t1-t0); walk_ns=(t2-t1)
print('kernel span ns', t2.max()-base)
print('tile start (ns after first): min/mean/max', (t0.min(1)-base).min(), (t0.min(1)-base).mean(), (t0.min(1)-base).max())
print('sort ns per tile mean/max', sort_ns.max(1).mean(), sort_ns.max(1).max())
print('walk ns per warp mean/median/p90/max', walk_ns.mean(), np.median(walk_ns), np.percentile(walk_ns,90), walk_ns.max())
print('tile total ns (max warp end - start) mean/max', (t2.max(1)-t0.min(1)).mean(), (t2.max(1)-t0.min(1)).max())
print('batches per warp mean/max', st[:,:,3].mean(), st[:,:,3].max(), ' survivors per warp mean/max', st[:,:,4].mean(), st[:,:,4].max())
tm=st[:,:,5]; tl=st[:,:,6]
print('mask wait ns mean/max', (tm-t1).mean(), (tm-t1).max(), ' loop ns mean/max', (tl-tm).mean(), (tl-tm).max(), ' shade ns mean/max', (t2-tl).mean(), (t2-tl).max())
end_by_sm={}
for t in range(300):
    sm=st[t,0,7]; end_by_sm.setdefault(sm,[]).append((t0[t].min()-base, t2[t].max()-base))
ends=sorted(max(e for _,e in v) for v in end_by_sm.values())
print('SMs used', len(end_by_sm), 'SM finish ns: min/median/max', ends[0], ends[len(ends)//2], ends[-1])
cnts=[len(v) for v in end_by_sm.values()]
print('tiles per SM histogram', np.bincount(cnts))
# correlation walk time vs survivors
w=walk_ns.reshape(-1); s=st[:,:,4].reshape(-1)
print('ns per survivor (fit)', np.polyfit(s,w,1))
worst=np.argsort(-(t2.max(1)-t0.min(1)))[:5]
for t in worst: print('tile',t,'total',(t2[t].max()-t0[t].min()),'sort',sort_ns[t].max(),'walk max',walk_ns[t].max(),'batches',st[t,:,3].max(),'surv',st[t,:,4].max())
