"""Level file -> room triangles through the reference's own compiled code (docs/bonnie-32.wasm):
`world::level::load_level_from_str` (RON parse + validation) and `Room::add_horizontal_face_to_render_data`,
`add_wall_to_render_data`, `add_diagonal_wall_to_render_data` (src/world/geometry.rs:2906-3352), driven in the order of
`Room::to_render_data_with_textures` (:2839-2904), which the compiler inlined into its callers.

TEST INFRASTRUCTURE (build container only).  Layouts recovered from the decompiled caller
`editor::layout::draw_player_camera_preview` (python wasmdecomp.py ... draw_player_camera_preview):

  Level            rooms Vec<Room>{cap@0, ptr@4, len@8}
  Room (116 B)     sectors Vec<Vec<Option<Sector>>>{ptr@4, len@8}; position x@68 y@72 z@76
  column (12 B)    Vec<Option<Sector>>{ptr@4, len@8}
  Option<Sector> (464 B)  i32@0 == 3: None; floor Option<HorizontalFace>@0 (i32 == 2: None); ceiling@196 (i32 == 2: None);
                   walls_north{ptr@396,len@400} east{408,412} south{420,424} west{432,436} nwse{444,448} nesw{456,460}
  VerticalFace     100 B
  add_horizontal_face_to_render_data(room_y f32, &mut Vec<Vertex>, &mut Vec<Face>, &face, base_x f32, base_z f32, grid_x, grid_z, is_floor, &&map)
  add_wall_to_render_data(room_y, verts, faces, &wall, base_x, base_z, grid_x, grid_z, direction 0..3 = N E S W, &map)
  add_diagonal_wall_to_render_data(room_y, verts, faces, &wall, base_x, base_z, grid_x, is_nwse, &map)
  The texture resolver of this monomorphisation is the editor's closure over a HashMap<(pack, name), id>; a map with
  items == 0 (32 zero bytes) makes every lookup miss, and the geometry code then uses its default (texture 0, width 64).
"""
import struct

import numpy as np

from ref_wasm import RefWasm

F_HORIZ, F_WALL, F_DIAG = 687, 688, 689


def room_geometry(w, room_ptr):
    """(vertices[n] as dict of arrays, faces) of one Room, in to_render_data_with_textures order."""
    rd = lambda a, n: w.read(a, n)
    u32 = lambda a: struct.unpack('<I', rd(a, 4))[0]
    f32 = lambda a: struct.unpack('<f', rd(a, 4))[0]
    verts = w.put(struct.pack('<III', 0, 4, 0), 4)
    faces = w.put(struct.pack('<III', 0, 4, 0), 4)
    hmap = w.put(b'\0' * 32, 8)
    hmap_ref = w.put(struct.pack('<I', hmap), 4)
    cols_ptr, ncols = u32(room_ptr + 4), u32(room_ptr + 8)
    rx, ry, rz = f32(room_ptr + 68), f32(room_ptr + 72), f32(room_ptr + 76)
    for gx in range(ncols):
        col = cols_ptr + gx * 12
        sp, ns = u32(col + 4), u32(col + 8)
        base_x = float(np.float32(np.float32(gx) * np.float32(1024.0)) + np.float32(rx))
        for gz in range(ns):
            s = sp + gz * 464
            if u32(s) == 3:
                continue
            base_z = float(np.float32(np.float32(gz) * np.float32(1024.0)) + np.float32(rz))
            if u32(s) != 2:
                w.call(F_HORIZ, ry, verts, faces, s, base_x, base_z, gx, gz, 1, hmap_ref)
            if u32(s + 196) != 2:
                w.call(F_HORIZ, ry, verts, faces, s + 196, base_x, base_z, gx, gz, 0, hmap_ref)
            for d, off in enumerate((396, 408, 420, 432)):
                wp, wn = u32(s + off), u32(s + off + 4)
                for k in range(wn):
                    w.call(F_WALL, ry, verts, faces, wp + k * 100, base_x, base_z, gx, gz, d, hmap)
            for nwse, off in ((1, 444), (0, 456)):
                wp, wn = u32(s + off), u32(s + off + 4)
                for k in range(wn):
                    w.call(F_DIAG, ry, verts, faces, wp + k * 100, base_x, base_z, gx, nwse, hmap)
    _, vp, vn = struct.unpack('<III', rd(verts, 12))
    _, fp, fn = struct.unpack('<III', rd(faces, 12))
    vb = np.frombuffer(rd(vp, vn * 44), np.uint8).reshape(vn, 44) if vn else np.zeros((0, 44), np.uint8)
    fb = np.frombuffer(rd(fp, fn * 24), np.uint8).reshape(fn, 24) if fn else np.zeros((0, 24), np.uint8)
    v = {"blend": vb[:, 8].copy(), "rgb": vb[:, 9:12].copy(), "pos": vb[:, 12:24].copy().view('<f4'),
         "uv": vb[:, 24:32].copy().view('<f4'), "normal": vb[:, 32:44].copy().view('<f4')}
    fw = fb[:, :20].copy().view('<u4') if fn else np.zeros((0, 5), np.uint32)
    f = {"tex_some": fw[:, 0], "tex": fw[:, 1], "v": fw[:, 2:5], "black_transparent": fb[:, 20].copy()}
    return v, f


def level_geometry(ron_text: bytes):
    """[(vertices, faces)] per room of a level (RON text, already brotli-decoded)."""
    w = RefWasm()
    p = w.put(ron_text, 1)
    out = w.alloc(1024, 8)
    w.write(out, b'\xAA' * 64)
    w.call('load_level_from_str', out, p, len(ron_text))
    cap, rooms_ptr, n_rooms = struct.unpack('<III', w.read(out, 12))
    assert n_rooms < 1000 and rooms_ptr > 4096, 'load_level_from_str failed (Err variant?)'
    return [room_geometry(w, rooms_ptr + i * 116) for i in range(n_rooms)]
