import sys, time, ctypes as C
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import __graft_entry__ as g
pkg=g.load_package()
import c3, cases
from bonnie32_b200.raster import Camera
ctx=pkg.Context(0)
kt=(C.c_float*7)()
for p in c3.scene_paths()[:3]:
    lv=c3.load_scene(p); ctx.set_textures(lv.textures)
    allv=np.concatenate([rc.vertices["pos"] for rc in lv.rooms]); lo,hi=allv.min(0),allv.max(0); ctr=(lo+hi)/2
    for name,(w,h) in (("320x240",(320,240)),("640x480",(640,480))):
        fb=pkg.Framebuffer(w,h,ctx)
        meshes=[pkg.Mesh(ctx, rc.vertices, rc.faces) for rc in lv.rooms]
        for cname, pos, ry in (("orbit", None, None), ("inside_centre", ctr, 0.0), ("inside_near_wall", np.array([lo[0]+0.1*(hi[0]-lo[0]), ctr[1], ctr[2]]), 1.57), ("inside_corner", lo+0.15*(hi-lo), 0.8)):
            if pos is None: cam=lv.camera
            else: cam=cases._rotated_camera(0.1, ry, pos.astype(np.float32))
            tot=np.zeros(2); drawn=0
            for r in range(8):
                fb.clear(lv.clear)
                for m,rc in zip(meshes, lv.rooms):
                    tm=m.render(cam, lv.settings(rc.ambient), rc.fog); 
                    ctx.lib.b32_debug_kernel_times(ctx.h, kt, 7)
                    if r>=3: tot+=np.array([kt[0],kt[1]]); drawn=tm["triangles_drawn"]
            tot/=5
            print(f"{lv.name:10s} {name} {cname:18s} drawn(last room)={drawn:5d} setup {tot[0]*1e3:6.1f} us fill {tot[1]*1e3:7.1f} us")
        for m in meshes: m.free()
