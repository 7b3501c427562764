"""Host submission cost vs device throughput of enqueued frames (GPU box only).
usage: python tools/hostbound.py [n_contexts] [n_mesh_copies]      env B32_NO_GRAPH=1 = plain launches"""
import sys, time, ctypes as C
sys.path.insert(0, '.')
import __graft_entry__ as g
pkg = g.load_package(); abi = pkg.abi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sc = pkg.scenes.scene_c4()
ctxs = [pkg.Context(0) for _ in range(N)]
fbs = [pkg.Framebuffer(320, 240, c) for c in ctxs]
for c in ctxs: c.set_textures(sc.textures)
meshes = [pkg.Mesh(ctxs[0], sc.vertices, sc.faces) for _ in range(M)]; ctxs[0].sync()
cam = sc.camera.to_abi(); st, keep = sc.settings.to_abi()
lib = ctxs[0].lib
clear4 = (C.c_uint8 * 4)(20, 22, 28, 255)
def step(k):
    c = ctxs[k % N]
    lib.b32_frame_15_enqueue(c.h, clear4, meshes[k % M].h, C.byref(cam), C.byref(st), None)
for k in range(4 * max(N, M) * 2): step(k)
for c in ctxs: c.sync()
for reps in (200, 200):
    t0 = time.perf_counter()
    for k in range(reps): step(k)
    t1 = time.perf_counter()
    for c in ctxs: c.sync()
    t2 = time.perf_counter()
    print(f"ctx={N} copies={M} reps={reps}: host submit {1e6*(t1-t0)/reps:.1f} us/frame, total {1e6*(t2-t0)/reps:.1f} us/frame, "
          f"graph launches {sum(c.graph_launches() for c in ctxs)}")
