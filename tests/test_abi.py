"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/b32_raster.h
declares, the POD layouts match the header, and the product has no CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import __graft_entry__ as entry
import bonnie32_b200 as pkg
from bonnie32_b200 import abi

ROOT = entry.ROOT
HEADER = os.path.join(ROOT, "include", "b32_raster.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b32_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    entry.build_cuda()
    lib = abi.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libb32raster.so does not export {n}"
    assert sorted(abi.SYMBOLS) == names, "abi.py and include/b32_raster.h disagree on the symbol list"


def test_pod_layouts_match_header(tmp_path):
    """sizeof/offsetof as the C compiler sees them == the ctypes / numpy mirrors."""
    prog = tmp_path / "sizes.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "b32_raster.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(b32_vertex), sizeof(b32_face), sizeof(b32_camera), sizeof(b32_light),
         sizeof(b32_settings), sizeof(b32_fog), sizeof(b32_timings), sizeof(b32_tex_desc));
  printf("%zu %zu %zu %zu %zu\n", offsetof(b32_vertex, uv), offsetof(b32_vertex, normal), offsetof(b32_vertex, r),
         offsetof(b32_settings, ambient), offsetof(b32_settings, lights));
  printf("%zu %zu %zu %zu\n", sizeof(b32_tex8_desc), offsetof(b32_tex8_desc, pixels), sizeof(b32_sky_vertex), offsetof(b32_sky_vertex, r));
  printf("%zu %zu %zu %zu %zu %d\n", sizeof(b32_line), offsetof(b32_line, z0), offsetof(b32_line, r), offsetof(b32_line, kind),
         offsetof(b32_line, alpha), B32_LINE_MAX_COORD);
  printf("%zu %zu\n", sizeof(b32_star), offsetof(b32_star, r));
  printf("%zu %zu\n", sizeof(b32_placement), offsetof(b32_placement, world_pos));
  return 0; }''')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes = [int(x) for x in out]
    assert sizes[:8] == [36, 16, 48, C.sizeof(abi.Light), C.sizeof(abi.Settings), C.sizeof(abi.Fog), C.sizeof(abi.Timings), C.sizeof(abi.TexDesc)]
    assert abi.VERTEX_DTYPE.itemsize == 36 and abi.FACE_DTYPE.itemsize == 16 and C.sizeof(abi.Camera) == 48
    assert sizes[8:13] == [12, 20, 32, abi.Settings.ambient.offset, abi.Settings.lights.offset]
    F = abi.LINE_DTYPE.fields
    assert sizes[23:25] == [abi.STAR_DTYPE.itemsize, abi.STAR_DTYPE.fields["rgb"][1]]
    assert sizes[25:] == [C.sizeof(abi.Placement), abi.Placement.world_pos.offset]
    assert sizes[17:23] == [abi.LINE_DTYPE.itemsize, F["z0"][1], F["rgb"][1], F["kind"][1], F["alpha"][1], abi.LINE_MAX_COORD]
    assert sizes[13:17] == [C.sizeof(abi.Tex8Desc), abi.Tex8Desc.pixels.offset, abi.SKY_VERTEX_DTYPE.itemsize, abi.SKY_VERTEX_DTYPE.fields["rgb"][1]]
    assert abi.VERTEX_DTYPE.fields["uv"][1] == 12 and abi.VERTEX_DTYPE.fields["normal"][1] == 20 and abi.VERTEX_DTYPE.fields["rgba"][1] == 32


def test_face_flags_packing():
    f = abi.face_flags(7, abi.BLEND_ADD, True, 200)
    assert int(f) == 7 | (2 << 16) | (1 << 19) | (200 << 24)
    assert int(abi.face_flags()) == 0xFFFF | (1 << 19) | (255 << 24)


def test_compact_marshalling_choice():
    """What a shim sends through b32_render_mesh_15_ex (abi.compact_buffers = rust/b32_shim.rs marshal_compact): normals are
    dropped only when nothing shades; an unindexed soup sends no indices; one flags word when all faces share it."""
    from bonnie32_b200 import scenes
    sc = scenes.scene_c4(n_tris=500)
    v, f, flags = abi.compact_buffers(sc.vertices, sc.faces, True)
    assert flags == abi.VTX_NO_NORMAL | abi.FACES_UNIFORM and v.dtype == abi.VERTEX_NN_DTYPE and f.shape == (1,) and f.dtype == np.uint32
    assert v.nbytes + f.nbytes == 500 * 3 * 24 + 4 and np.array_equal(v["pos"], sc.vertices["pos"]) and f[0] == sc.faces["flags"][0]
    v, f, flags = abi.compact_buffers(sc.vertices, sc.faces, False)          # something shades: the normals travel
    assert flags == abi.FACES_UNIFORM and v.dtype == abi.VERTEX_DTYPE
    f2 = sc.faces.copy(); f2["flags"][7] = abi.face_flags(0, abi.BLEND_ADD, True, 255)
    v, f, flags = abi.compact_buffers(sc.vertices, f2, True)                 # mixed flags: one word per face
    assert flags == abi.VTX_NO_NORMAL | abi.FACES_IMPLICIT and f.shape == (500,) and np.array_equal(f, f2["flags"])
    f3 = sc.faces.copy(); f3["v"][3] = f3["v"][3][::-1]
    v, f, flags = abi.compact_buffers(sc.vertices, f3, True)                 # not a soup in order: the full face records
    assert flags == abi.VTX_NO_NORMAL and f.dtype == abi.FACE_DTYPE and len(f) == 500


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device context creation fails loudly; nothing renders on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.B32Error) as e:
        pkg.Context(0)
    assert e.value.code == abi.B32_ERR_NO_DEVICE


def test_missing_library_fails_loudly(tmp_path):
    saved = abi._lib
    abi._lib = None
    try:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            abi.load_library(str(tmp_path / "nope.so"))
    finally:
        abi._lib = saved


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under bonnie-32_b200/ or include/ may reference it."""
    bad = []
    for base in (entry.PKG_DIR, os.path.join(ROOT, "include")):
        for dp, _, files in os.walk(base):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"\boracle\b|b32o_|pymodel", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """Outside oracle/ itself, only tests/, __graft_entry__ (build + smoke) and bench.py (its CPU legs) may import or load
    anything of the oracle: the measurement tools and the Rust shim must not."""
    bad = []
    for base in ("tools", "rust"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"from oracle|import oracle|libb32oracle|b32o_|pymodel", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
    # bench.py: the oracle is loaded inside time_oracle / _oracle_worker only (cpu_baseline and --impl reference)
    src = open(os.path.join(ROOT, "bench.py")).read()
    for m in re.finditer(r"from oracle import|import oracle", src):
        head = src[:m.start()]
        fn = re.findall(r"^def (\w+)", head, flags=re.M)[-1]
        assert fn in ("time_oracle", "_oracle_worker", "_oracle_worker_init"), fn


def test_built_for_sm100a_with_exact_arithmetic_flags():
    """The cubin targets sm_100a and is compiled without FMA contraction / with IEEE div+sqrt, no FTZ.
    (FFMA still appears in SASS inside the correctly-rounded division sequences, so the flags — and
    the GPU parity tests — are what is checked.)"""
    for flag in ("-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "arch=compute_100a,code=sm_100a", "-lineinfo"):
        assert flag in entry.NVCC_FLAGS
    entry.build_cuda()
    out = subprocess.run(["cuobjdump", "-lelf", entry.LIB], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_cpp_mirror_compiles_and_links(tmp_path):
    """include/b32_raster.hpp (the C++ host mirror) compiles with -Wall -Wextra -Werror against the header and links
    against the built library: every entry point it wraps exists."""
    src = tmp_path / "mirror.cpp"
    src.write_text(r'''
#include "b32_raster.hpp"
// instantiate every wrapper without running anything (no GPU here)
int use(b32::Context& c) {
    b32::Framebuffer fb(c, 4, 4);
    const uint8_t a[3] = {1, 2, 3}, b[3] = {4, 5, 6};
    fb.clear_gradient(a, b);
    b32_camera cam{}; b32_settings st{};
    fb.render_skybox_mesh({}, {}, cam); fb.render_stars({}, cam, 2.0f); fb.draw_lines({});
    std::vector<b32_vertex> v; std::vector<b32_face> f;
    b32::render_mesh_15(fb, v, f, cam, st); b32::render_mesh(fb, v, f, cam, st);
    b32::Mesh m(c, v, f);
    const float wp[3] = {0, 0, 0};
    m.render_15(cam, st); m.frame_15_enqueue(nullptr, cam, st); m.render_placed(0.5f, 0.87f, 0.48f, wp, cam, st);
    c.set_textures({}); c.set_textures_rgb888({}); c.sync();
    return (int)fb.pixels().size() + (int)fb.zbuffer().size();
}
int main(int argc, char**) { if (argc > 99) { b32::Context c(0); return use(c); } return 0; }
''')
    exe = tmp_path / "mirror"
    lib_dir = os.path.dirname(abi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                           "-L", lib_dir, "-l:" + os.path.basename(abi.LIB_PATH), "-Wl,-rpath," + lib_dir])
    subprocess.check_call([str(exe)])
