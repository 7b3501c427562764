"""Text disassembler for single functions of a WebAssembly MVP(+sign-ext, sat-trunc, bulk-memory) module.

Test infrastructure: used once, by hand, to recover the Rust struct layouts the reference's compiled
rasterizer functions expect (recorded in ref_wasm.py).  Usage:
    python wasmdis.py /root/reference/docs/bonnie-32.wasm render_mesh_15
"""
import struct
import sys

from wasmparse import Module, leb_u, leb_s

SIMPLE = {
    0x00: 'unreachable', 0x01: 'nop', 0x05: 'else', 0x0B: 'end', 0x0F: 'return', 0x1A: 'drop', 0x1B: 'select',
    0x45: 'i32.eqz', 0x46: 'i32.eq', 0x47: 'i32.ne', 0x48: 'i32.lt_s', 0x49: 'i32.lt_u', 0x4A: 'i32.gt_s',
    0x4B: 'i32.gt_u', 0x4C: 'i32.le_s', 0x4D: 'i32.le_u', 0x4E: 'i32.ge_s', 0x4F: 'i32.ge_u',
    0x50: 'i64.eqz', 0x51: 'i64.eq', 0x52: 'i64.ne', 0x53: 'i64.lt_s', 0x54: 'i64.lt_u', 0x55: 'i64.gt_s',
    0x56: 'i64.gt_u', 0x57: 'i64.le_s', 0x58: 'i64.le_u', 0x59: 'i64.ge_s', 0x5A: 'i64.ge_u',
    0x5B: 'f32.eq', 0x5C: 'f32.ne', 0x5D: 'f32.lt', 0x5E: 'f32.gt', 0x5F: 'f32.le', 0x60: 'f32.ge',
    0x61: 'f64.eq', 0x62: 'f64.ne', 0x63: 'f64.lt', 0x64: 'f64.gt', 0x65: 'f64.le', 0x66: 'f64.ge',
    0x67: 'i32.clz', 0x68: 'i32.ctz', 0x69: 'i32.popcnt', 0x6A: 'i32.add', 0x6B: 'i32.sub', 0x6C: 'i32.mul',
    0x6D: 'i32.div_s', 0x6E: 'i32.div_u', 0x6F: 'i32.rem_s', 0x70: 'i32.rem_u', 0x71: 'i32.and', 0x72: 'i32.or',
    0x73: 'i32.xor', 0x74: 'i32.shl', 0x75: 'i32.shr_s', 0x76: 'i32.shr_u', 0x77: 'i32.rotl', 0x78: 'i32.rotr',
    0x79: 'i64.clz', 0x7A: 'i64.ctz', 0x7B: 'i64.popcnt', 0x7C: 'i64.add', 0x7D: 'i64.sub', 0x7E: 'i64.mul',
    0x7F: 'i64.div_s', 0x80: 'i64.div_u', 0x81: 'i64.rem_s', 0x82: 'i64.rem_u', 0x83: 'i64.and', 0x84: 'i64.or',
    0x85: 'i64.xor', 0x86: 'i64.shl', 0x87: 'i64.shr_s', 0x88: 'i64.shr_u', 0x89: 'i64.rotl', 0x8A: 'i64.rotr',
    0x8B: 'f32.abs', 0x8C: 'f32.neg', 0x8D: 'f32.ceil', 0x8E: 'f32.floor', 0x8F: 'f32.trunc', 0x90: 'f32.nearest',
    0x91: 'f32.sqrt', 0x92: 'f32.add', 0x93: 'f32.sub', 0x94: 'f32.mul', 0x95: 'f32.div', 0x96: 'f32.min',
    0x97: 'f32.max', 0x98: 'f32.copysign',
    0x99: 'f64.abs', 0x9A: 'f64.neg', 0x9B: 'f64.ceil', 0x9C: 'f64.floor', 0x9D: 'f64.trunc', 0x9E: 'f64.nearest',
    0x9F: 'f64.sqrt', 0xA0: 'f64.add', 0xA1: 'f64.sub', 0xA2: 'f64.mul', 0xA3: 'f64.div', 0xA4: 'f64.min',
    0xA5: 'f64.max', 0xA6: 'f64.copysign',
    0xA7: 'i32.wrap_i64', 0xA8: 'i32.trunc_f32_s', 0xA9: 'i32.trunc_f32_u', 0xAA: 'i32.trunc_f64_s',
    0xAB: 'i32.trunc_f64_u', 0xAC: 'i64.extend_i32_s', 0xAD: 'i64.extend_i32_u', 0xAE: 'i64.trunc_f32_s',
    0xAF: 'i64.trunc_f32_u', 0xB0: 'i64.trunc_f64_s', 0xB1: 'i64.trunc_f64_u', 0xB2: 'f32.convert_i32_s',
    0xB3: 'f32.convert_i32_u', 0xB4: 'f32.convert_i64_s', 0xB5: 'f32.convert_i64_u', 0xB6: 'f32.demote_f64',
    0xB7: 'f64.convert_i32_s', 0xB8: 'f64.convert_i32_u', 0xB9: 'f64.convert_i64_s', 0xBA: 'f64.convert_i64_u',
    0xBB: 'f64.promote_f32', 0xBC: 'i32.reinterpret_f32', 0xBD: 'i64.reinterpret_f64',
    0xBE: 'f32.reinterpret_i32', 0xBF: 'f64.reinterpret_i64',
    0xC0: 'i32.extend8_s', 0xC1: 'i32.extend16_s', 0xC2: 'i64.extend8_s', 0xC3: 'i64.extend16_s',
    0xC4: 'i64.extend32_s',
}
MEM = {
    0x28: 'i32.load', 0x29: 'i64.load', 0x2A: 'f32.load', 0x2B: 'f64.load', 0x2C: 'i32.load8_s',
    0x2D: 'i32.load8_u', 0x2E: 'i32.load16_s', 0x2F: 'i32.load16_u', 0x30: 'i64.load8_s', 0x31: 'i64.load8_u',
    0x32: 'i64.load16_s', 0x33: 'i64.load16_u', 0x34: 'i64.load32_s', 0x35: 'i64.load32_u',
    0x36: 'i32.store', 0x37: 'i64.store', 0x38: 'f32.store', 0x39: 'f64.store', 0x3A: 'i32.store8',
    0x3B: 'i32.store16', 0x3C: 'i64.store8', 0x3D: 'i64.store16', 0x3E: 'i64.store32',
}
FC = {0: 'i32.trunc_sat_f32_s', 1: 'i32.trunc_sat_f32_u', 2: 'i32.trunc_sat_f64_s', 3: 'i32.trunc_sat_f64_u',
      4: 'i64.trunc_sat_f32_s', 5: 'i64.trunc_sat_f32_u', 6: 'i64.trunc_sat_f64_s', 7: 'i64.trunc_sat_f64_u',
      10: 'memory.copy', 11: 'memory.fill'}
VT = {0x7F: 'i32', 0x7E: 'i64', 0x7D: 'f32', 0x7C: 'f64', 0x40: ''}


def func_body(m, idx):
    d = m.data
    p, e = m.sections[10]
    n, p = leb_u(d, p)
    k = idx - m.n_func_imports
    for i in range(n):
        sz, p = leb_u(d, p)
        if i == k:
            return p, p + sz
        p += sz
    raise IndexError(idx)


def disasm(m, idx, out=sys.stdout):
    d = m.data
    p, e = func_body(m, idx)
    params, res = m.sig(idx)
    print(f'func {idx} {m.names.get(idx)} params={[VT[t] for t in params]} results={[VT[t] for t in res]}', file=out)
    ng, p = leb_u(d, p)
    li = len(params)
    for _ in range(ng):
        c, p = leb_u(d, p)
        t = d[p]
        p += 1
        print(f'  locals {li}..{li + c - 1}: {VT[t]}', file=out)
        li += c
    depth = 0
    while p < e:
        at = p
        op = d[p]
        p += 1
        ind = '  ' * (depth + 1)
        if op in (0x02, 0x03, 0x04):
            bt, p = leb_s(d, p)
            nm = {2: 'block', 3: 'loop', 4: 'if'}[op]
            print(f'{at:7d}{ind}{nm} {bt if bt >= 0 else VT.get(bt & 0x7F, bt)}  ;; @{depth}', file=out)
            depth += 1
        elif op == 0x05:
            print(f'{at:7d}{"  " * depth}else', file=out)
        elif op == 0x0B:
            depth -= 1
            print(f'{at:7d}{"  " * (depth + 1)}end', file=out)
        elif op in (0x0C, 0x0D):
            l, p = leb_u(d, p)
            print(f'{at:7d}{ind}{"br" if op == 0x0C else "br_if"} {l}', file=out)
        elif op == 0x0E:
            n, p = leb_u(d, p)
            ls = []
            for _ in range(n + 1):
                l, p = leb_u(d, p)
                ls.append(l)
            print(f'{at:7d}{ind}br_table {ls}', file=out)
        elif op == 0x10:
            f, p = leb_u(d, p)
            print(f'{at:7d}{ind}call {f} <{m.names.get(f)}>', file=out)
        elif op == 0x11:
            t, p = leb_u(d, p)
            tb, p = leb_u(d, p)
            print(f'{at:7d}{ind}call_indirect type={t}', file=out)
        elif op in (0x20, 0x21, 0x22, 0x23, 0x24):
            x, p = leb_u(d, p)
            nm = {0x20: 'local.get', 0x21: 'local.set', 0x22: 'local.tee', 0x23: 'global.get', 0x24: 'global.set'}[op]
            print(f'{at:7d}{ind}{nm} {x}', file=out)
        elif op in MEM:
            a, p = leb_u(d, p)
            o, p = leb_u(d, p)
            print(f'{at:7d}{ind}{MEM[op]} offset={o}', file=out)
        elif op in (0x3F, 0x40):
            p += 1
            print(f'{at:7d}{ind}{"memory.size" if op == 0x3F else "memory.grow"}', file=out)
        elif op == 0x41:
            v, p = leb_s(d, p)
            print(f'{at:7d}{ind}i32.const {v}', file=out)
        elif op == 0x42:
            v, p = leb_s(d, p)
            print(f'{at:7d}{ind}i64.const {v}', file=out)
        elif op == 0x43:
            v = struct.unpack_from('<f', d, p)[0]
            raw = struct.unpack_from('<I', d, p)[0]
            p += 4
            print(f'{at:7d}{ind}f32.const {v!r} (0x{raw:08x})', file=out)
        elif op == 0x44:
            v = struct.unpack_from('<d', d, p)[0]
            p += 8
            print(f'{at:7d}{ind}f64.const {v!r}', file=out)
        elif op == 0xFC:
            s, p = leb_u(d, p)
            if s == 10:
                p += 2
            elif s == 11:
                p += 1
            print(f'{at:7d}{ind}{FC.get(s, f"fc.{s}")}', file=out)
        elif op == 0x1C:
            n, p = leb_u(d, p)
            p += n
            print(f'{at:7d}{ind}select_t', file=out)
        elif op in SIMPLE:
            print(f'{at:7d}{ind}{SIMPLE[op]}', file=out)
        else:
            print(f'{at:7d}{ind}?? 0x{op:02x}', file=out)
            break


if __name__ == '__main__':
    m = Module(sys.argv[1])
    for key in sys.argv[2:]:
        if key.isdigit():
            disasm(m, int(key))
        else:
            for i, n in m.find(key):
                disasm(m, i)
