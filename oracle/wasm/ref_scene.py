"""Runs the reference's compiled `render_mesh_15` / `render_mesh` (docs/bonnie-32.wasm) on a `scenes.Scene`.

TEST INFRASTRUCTURE (oracle side, build container only).

Struct layouts of the wasm32 build (rustc 1.92 reorders fields; recovered with RefWasm.watch() load traces and by
calling the binary's own constructors `Framebuffer::new`, `RasterSettings::game` — see LAYOUTS.md):

  Framebuffer (32 B)   pixels Vec<u8>{cap,ptr,len}@0  zbuffer Vec<f32>{cap,ptr,len}@12  width@24  height@28
  Vertex (44 B)        bone_index Option<usize>@0  color{blend@8,r@9,g@10,b@11}  pos@12  uv@24  normal@32
  Face (24 B)          texture_id Option<usize>{tag@0,val@4}  v0@8 v1@12 v2@16  black_transparent@20
                       (0.1.8 has no Face.blend_mode / editor_alpha: blend modes come in a separate
                        `face_blend_modes: Option<&[BlendMode]>` argument)
  Texture15 (36 B)     pixels Vec<u16>{cap,ptr,len}@0  name String@12  width@24  height@28  blend_mode@32
  Camera (56 B)        position@0  rotation_x@12  rotation_y@16  basis_x@20  basis_y@32  basis_z@44
  RasterSettings (44B) ortho Option{tag@0,zoom@4,cx@8,cy@12}  lights Vec{cap,ptr,len}@16  ambient@28
                       affine@32 zbuffer@33 cull@34 bf_wire@35 lowres@36 dither@37 stretch@38 wire_overlay@39
                       rgb555@40 fixed@41 xray@42 shading@43
  Light (60 B)         type tag@0 (0 Directional{dir@4}, 1 Point{pos@4,radius@16}, 2 Spot{pos@4,dir@16,angle@28,radius@32})  name String@36
                       color{blend@48,r@49,g@50,b@51}  intensity@52  enabled@56
  fog Option<(f32,f32,f32,Color)> (16 B, by pointer)  start@0 falloff@4 cull@8 color{blend@12 (6 = None),r,g,b}
  RasterTimings (32 B) six f32 + triangles_drawn u32 @24
"""
import struct

import numpy as np

from ref_wasm import RefWasm, enable_get_time

DANGLING = 4


class RefRasterizer:
    def __init__(self):
        self.w = RefWasm()
        enable_get_time(self.w)

    # ---- marshalling --------------------------------------------------------------------------
    def _vertices(self, v):
        n = len(v)
        b = np.zeros((n, 44), np.uint8)
        rgba = np.asarray(v['rgba'], np.uint8)
        b[:, 8] = rgba[:, 3]
        b[:, 9:12] = rgba[:, :3]
        b[:, 12:24] = np.ascontiguousarray(v['pos'], '<f4').view(np.uint8).reshape(n, 12)
        b[:, 24:32] = np.ascontiguousarray(v['uv'], '<f4').view(np.uint8).reshape(n, 8)
        b[:, 32:44] = np.ascontiguousarray(v['normal'], '<f4').view(np.uint8).reshape(n, 12)
        return b

    def _faces(self, f):
        n = len(f)
        flags = np.asarray(f['flags'], np.uint32)
        tex = flags & 0xFFFF
        rec = np.zeros((n, 6), np.uint32)
        rec[:, 0] = (tex != 0xFFFF).astype(np.uint32)
        rec[:, 1] = np.where(tex != 0xFFFF, tex, 0)
        rec[:, 2:5] = f['v']
        rec[:, 5] = (flags >> 19) & 1
        blend = ((flags >> 16) & 7).astype(np.uint8)
        alpha = (flags >> 24).astype(np.uint8)
        return rec, blend, alpha

    def _put(self, arr, align=4):
        b = arr.tobytes() if isinstance(arr, np.ndarray) else bytes(arr)
        if not b:
            return DANGLING
        p = self.w.alloc(len(b), align)
        self.w.write(p, b)
        self._allocs.append((p, len(b), align))
        return p

    def _settings(self, s):
        w = self.w
        lights = b''
        for l in s.lights:
            name = self._put(b'L')
            tag = int(l.type)
            payload = [0.0] * 8
            if tag == 0:
                payload[0:3] = [float(x) for x in l.direction]
            elif tag == 1:
                payload[0:3] = [float(x) for x in l.position]
                payload[3] = float(l.radius)
            else:                     # Spot{position@4, direction@16, angle@28, radius@32}: read off shade_multi_light_color's loads
                payload[0:3] = [float(x) for x in l.position]
                payload[3:6] = [float(x) for x in l.direction]
                payload[6] = float(l.angle)
                payload[7] = float(l.radius)
            lights += struct.pack('<I8f', tag, *payload) + struct.pack('<III', 1, name, 1) \
                + bytes([0, l.color[0], l.color[1], l.color[2]]) + struct.pack('<f', l.intensity) \
                + bytes([1 if l.enabled else 0, 0, 0, 0])
        lp = self._put(lights) if lights else DANGLING
        o = s.ortho_projection
        b = struct.pack('<Ifff', 1 if o else 0, *(o if o else (0.0, 0.0, 0.0)))
        b += struct.pack('<IIIf', len(s.lights), lp, len(s.lights), s.ambient)
        b += bytes([s.affine_textures, s.use_zbuffer, s.backface_cull, s.backface_wireframe, 0, s.dithering, 1,
                    s.wireframe_overlay, s.use_rgb555, s.use_fixed_point, s.xray_mode, int(s.shading)])
        assert len(b) == 44
        return self._put(b)

    def _camera(self, c):
        b = struct.pack('<3f2f3f3f3f', *[float(x) for x in c.position], float(getattr(c, 'rotation_x', 0.0)),
                        float(getattr(c, 'rotation_y', 0.0)), *[float(x) for x in c.basis_x],
                        *[float(x) for x in c.basis_y], *[float(x) for x in c.basis_z])
        return self._put(b)

    def _textures15(self, textures, expand):
        recs = b''
        for t in textures:
            t = expand(t)
            px = np.ascontiguousarray(t.pixels, '<u2').reshape(-1)
            pp = self._put(px, 2)
            nm = self._put(b'T')
            recs += struct.pack('<IIIIIIIII', len(px), pp, len(px), 1, nm, 1, t.width, t.height, int(t.blend_mode))
        return self._put(recs) if recs else DANGLING

    def _textures8(self, textures):
        """RGB888 `Texture` (same 36-byte layout as Texture15); a texel is a `Color` {blend@0, r@1, g@2, b@3}."""
        recs = b''
        for t in textures:
            px = np.ascontiguousarray(t.pixels, np.uint8).reshape(-1, 4)          # host order r, g, b, blend
            w = np.empty_like(px)
            w[:, 0] = px[:, 3]
            w[:, 1:4] = px[:, 0:3]
            pp = self._put(w, 1)
            nm = self._put(b'T')
            recs += struct.pack('<IIIIIIIII', len(px), pp, len(px), 1, nm, 1, t.width, t.height, int(t.blend_mode))
        return self._put(recs) if recs else DANGLING

    # ---- calls -------------------------------------------------------------------------------------
    def new_framebuffer(self, width, height, rgba=None, z=None):
        w = self.w
        fb = w.alloc(32, 4)
        w.call('Framebuffer3new', fb, width, height)
        if rgba is not None:
            ptr = struct.unpack('<I', w.read(fb + 4, 4))[0]
            w.write(ptr, np.ascontiguousarray(rgba, np.uint8))
        if z is not None:
            ptr = struct.unpack('<I', w.read(fb + 16, 4))[0]
            w.write(ptr, np.ascontiguousarray(z, '<f4'))
        return fb

    def read_framebuffer(self, fb):
        w = self.w
        _, pp, pl, _, zp, zl, width, height = struct.unpack('<8I', w.read(fb, 32))
        rgba = np.frombuffer(w.read(pp, pl), np.uint8).reshape(height, width, 4).copy()
        z = np.frombuffer(w.read(zp, zl * 4), '<f4').reshape(height, width).copy()
        return rgba, z

    def free_framebuffer(self, fb):
        w = self.w
        pc, pp, _, zc, zp, _, _, _ = struct.unpack('<8I', w.read(fb, 32))
        if pc:
            w.free(pp, pc, 1)
        if zc:
            w.free(zp, zc * 4, 4)
        w.free(fb, 32, 4)

    def render_mesh_15(self, fb, vertices, faces, textures, camera, settings, fog=None, expand=lambda t: t):
        """Returns triangles_drawn.  Face blend modes travel in 0.1.8's `face_blend_modes` slice; editor_alpha
        does not exist in 0.1.8 (must be 255)."""
        w = self.w
        self._allocs = []
        vb = self._vertices(vertices)
        rec, blend, alpha = self._faces(faces)
        assert (alpha == 255).all(), 'editor_alpha is not in the 0.1.8 binary'
        vp = self._put(vb)
        fp = self._put(rec)
        bp = self._put(blend, 1) if blend.any() else 0
        tp = self._textures15(textures, expand)
        cp = self._camera(camera)
        sp = self._settings(settings)
        if fog is None:
            fg = struct.pack('<fffBBBB', 0, 0, 0, 6, 0, 0, 0)
        else:
            start, falloff, cull, col = fog
            fg = struct.pack('<fffBBBB', start, falloff, cull, 0, col[0], col[1], col[2])
        fgp = self._put(fg)
        tim = self._put(b'\0' * 32, 8)
        w.call('render_mesh_15', tim, fb, vp, len(vertices), fp, len(faces), tp, len(textures), bp,
               len(faces) if bp else 0, cp, sp, fgp)
        drawn = struct.unpack('<I', w.read(tim + 24, 4))[0]
        for p, n, a in reversed(self._allocs):
            w.free(p, n, a)
        return drawn

    def render_mesh(self, fb, vertices, faces, textures8, camera, settings):
        """The RGB888 sibling (render.rs:1971-2259).  0.1.8 has neither Face.blend_mode nor editor_alpha here either."""
        w = self.w
        self._allocs = []
        rec, blend, alpha = self._faces(faces)
        assert (alpha == 255).all(), 'editor_alpha is not in the 0.1.8 binary'
        vp = self._put(self._vertices(vertices))
        fp = self._put(rec)
        tp = self._textures8(textures8)
        cp = self._camera(camera)
        sp = self._settings(settings)
        tim = self._put(b'\0' * 32, 8)
        w.call('render11render_mesh17', tim, fb, vp, len(vertices), fp, len(faces), tp, len(textures8), cp, sp)
        drawn = struct.unpack('<I', w.read(tim + 24, 4))[0]
        for p, n, a in reversed(self._allocs):
            w.free(p, n, a)
        return drawn

    def render_scene888(self, scene):
        r, g, b = scene.clear[:3]
        rgba = np.empty((scene.height, scene.width, 4), np.uint8)
        rgba[:] = (r, g, b, 255)
        z = np.full((scene.height, scene.width), np.finfo(np.float32).max, np.float32)
        fb = self.new_framebuffer(scene.width, scene.height, rgba, z)
        drawn = self.render_mesh(fb, scene.vertices, scene.faces, scene.textures8, scene.camera, scene.settings)
        out = self.read_framebuffer(fb)
        self.free_framebuffer(fb)
        return out[0], out[1], drawn

    def render_scene(self, scene, expand=lambda t: t):
        """Clear to scene.clear + render_mesh_15.  Returns (rgba, z, triangles_drawn)."""
        r, g, b = scene.clear[:3]
        rgba = np.empty((scene.height, scene.width, 4), np.uint8)
        rgba[:] = (r, g, b, 255)
        z = np.full((scene.height, scene.width), np.finfo(np.float32).max, np.float32)
        fb = self.new_framebuffer(scene.width, scene.height, rgba, z)
        drawn = self.render_mesh_15(fb, scene.vertices, scene.faces, scene.textures, scene.camera, scene.settings,
                                    scene.fog, expand)
        out = self.read_framebuffer(fb)
        self.free_framebuffer(fb)
        return out[0], out[1], drawn
