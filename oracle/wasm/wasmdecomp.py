"""Pseudo-code listing of one wasm function: folds the operand stack into expressions so the reference's
compiled rasterizer can be READ (layout / behaviour recovery for ref_scene.py).  Test infrastructure.

    python wasmdecomp.py /root/reference/docs/bonnie-32.wasm rasterize_triangle_15 > /tmp/rt15.c
"""
import re
import struct
import sys

from wasmparse import Module, leb_u, leb_s
from wasmdis import SIMPLE, MEM, FC, func_body

BINOPS = {'add': '+', 'sub': '-', 'mul': '*', 'div': '/', 'div_s': '/s', 'div_u': '/u', 'rem_s': '%s', 'rem_u': '%u',
          'and': '&', 'or': '|', 'xor': '^', 'shl': '<<', 'shr_s': '>>s', 'shr_u': '>>u', 'eq': '==', 'ne': '!=',
          'lt': '<', 'gt': '>', 'le': '<=', 'ge': '>=', 'lt_s': '<s', 'lt_u': '<u', 'gt_s': '>s', 'gt_u': '>u',
          'le_s': '<=s', 'le_u': '<=u', 'ge_s': '>=s', 'ge_u': '>=u'}
FUNCS2 = {'min', 'max', 'copysign', 'rotl', 'rotr'}


def short(name):
    if not name:
        return '?'
    m = re.findall(r'\d+([A-Za-z_][A-Za-z0-9_]*)', name)
    return '::'.join(m[-3:-1]) if len(m) >= 3 else name[:60]


def decomp(m, idx, out=sys.stdout):
    d = m.data
    p, e = func_body(m, idx)
    params, res = m.sig(idx)
    print(f'// func {idx} {m.names.get(idx)}  params={len(params)} results={len(res)}', file=out)
    ng, p = leb_u(d, p)
    for _ in range(ng):
        c, p = leb_u(d, p)
        p += 1
    stack = []
    depth = 0
    tmp = [0]

    def emit(s):
        print(f'{"  " * depth}{s}', file=out)

    def spill(token=None):
        for i, s in enumerate(stack):
            if token is None or re.search(r'\b%s\b' % re.escape(token), s) or (token == 'MEM' and '[' in s):
                if re.fullmatch(r't\d+|-?\d+|-?[\d.e+-]+f?|l\d+', s) and token is None:
                    continue
                t = f't{tmp[0]}'
                tmp[0] += 1
                emit(f'{t} = {s}')
                stack[i] = t

    def pop():
        return stack.pop() if stack else '<?>'

    while p < e:
        op = d[p]
        p += 1
        if op in (0x02, 0x03, 0x04):
            bt, p = leb_s(d, p)
            if op == 0x04:
                c = pop()
                spill()
                emit(f'if ({c}) {{  // @{depth}')
            else:
                spill()
                emit(f'{"block" if op == 2 else "loop"} {{  // @{depth}')
            depth += 1
        elif op == 0x05:
            spill()
            depth -= 1
            emit('} else {')
            depth += 1
        elif op == 0x0B:
            spill()
            depth -= 1
            if depth >= 0:
                emit('}')
        elif op == 0x0C:
            l, p = leb_u(d, p)
            spill()
            emit(f'br {l}  // -> @{depth - 1 - l}')
        elif op == 0x0D:
            l, p = leb_u(d, p)
            c = pop()
            spill()
            emit(f'if ({c}) br {l}  // -> @{depth - 1 - l}')
        elif op == 0x0E:
            n, p = leb_u(d, p)
            ls = []
            for _ in range(n + 1):
                l, p = leb_u(d, p)
                ls.append(l)
            c = pop()
            emit(f'br_table {c} {ls}  // depth now {depth}')
        elif op == 0x0F:
            emit(f'return {" ".join(stack[-len(res):]) if res else ""}')
        elif op == 0x10:
            f, p = leb_u(d, p)
            pr, rs = m.sig(f)
            args = [pop() for _ in pr][::-1]
            call = f'{short(m.names.get(f))}#{f}({", ".join(args)})'
            spill('MEM')
            if rs:
                t = f't{tmp[0]}'
                tmp[0] += 1
                emit(f'{t} = {call}')
                stack.append(t)
            else:
                emit(call)
        elif op == 0x11:
            t, p = leb_u(d, p)
            _, p = leb_u(d, p)
            pr, rs = m.types[t]
            fi = pop()
            args = [pop() for _ in pr][::-1]
            call = f'indirect[{fi}]({", ".join(args)})'
            spill('MEM')
            if rs:
                tt = f't{tmp[0]}'
                tmp[0] += 1
                emit(f'{tt} = {call}')
                stack.append(tt)
            else:
                emit(call)
        elif op == 0x1A:
            pop()
        elif op in (0x1B, 0x1C):
            if op == 0x1C:
                n, p = leb_u(d, p)
                p += n
            c = pop()
            b = pop()
            a = pop()
            stack.append(f'({c} ? {a} : {b})')
        elif op == 0x20:
            x, p = leb_u(d, p)
            stack.append(f'l{x}')
        elif op == 0x21:
            x, p = leb_u(d, p)
            v = pop()
            spill(f'l{x}')
            emit(f'l{x} = {v}')
        elif op == 0x22:
            x, p = leb_u(d, p)
            v = pop()
            spill(f'l{x}')
            emit(f'l{x} = {v}')
            stack.append(f'l{x}')
        elif op == 0x23:
            x, p = leb_u(d, p)
            stack.append(f'g{x}')
        elif op == 0x24:
            x, p = leb_u(d, p)
            emit(f'g{x} = {pop()}')
        elif op in MEM:
            _, p = leb_u(d, p)
            o, p = leb_u(d, p)
            nm = MEM[op]
            ty = nm.replace('.load', '').replace('.store', '')
            if 'load' in nm:
                a = pop()
                stack.append(f'{ty}[{a}+{o}]' if o else f'{ty}[{a}]')
            else:
                v = pop()
                a = pop()
                spill('MEM')
                emit(f'{ty}[{a}+{o}] = {v}' if o else f'{ty}[{a}] = {v}')
        elif op == 0x3F:
            p += 1
            stack.append('memory.size')
        elif op == 0x40:
            p += 1
            stack.append(f'memory.grow({pop()})')
        elif op == 0x41:
            v, p = leb_s(d, p)
            stack.append(str(v))
        elif op == 0x42:
            v, p = leb_s(d, p)
            stack.append(f'{v}L')
        elif op == 0x43:
            v = struct.unpack_from('<f', d, p)[0]
            p += 4
            stack.append(f'{v!r}f')
        elif op == 0x44:
            v = struct.unpack_from('<d', d, p)[0]
            p += 8
            stack.append(f'{v!r}')
        elif op == 0xFC:
            s, p = leb_u(d, p)
            if s == 10:
                p += 2
                n = pop(); sr = pop(); ds = pop()
                spill('MEM')
                emit(f'memcpy({ds}, {sr}, {n})')
            elif s == 11:
                p += 1
                n = pop(); v = pop(); ds = pop()
                spill('MEM')
                emit(f'memset({ds}, {v}, {n})')
            else:
                stack.append(f'{FC[s]}({pop()})')
        elif op in SIMPLE:
            nm = SIMPLE[op]
            if nm in ('unreachable',):
                emit('unreachable')
                continue
            if nm == 'nop':
                continue
            ty, o = nm.split('.')
            if o in BINOPS:
                b = pop()
                a = pop()
                pre = 'f:' if ty[0] == 'f' and o in ('add', 'sub', 'mul', 'div') else ''
                l64 = 'L' if ty == 'i64' else ''
                stack.append(f'({a} {BINOPS[o]}{l64} {b})')
            elif o in FUNCS2:
                b = pop()
                a = pop()
                stack.append(f'{o}({a}, {b})')
            elif o == 'eqz':
                stack.append(f'!({pop()})')
            else:
                stack.append(f'{nm}({pop()})')
        else:
            emit(f'?? 0x{op:02x}')
            break


if __name__ == '__main__':
    m = Module(sys.argv[1])
    for key in sys.argv[2:]:
        if key.isdigit():
            decomp(m, int(key))
        else:
            for i, n in m.find(key):
                decomp(m, i)
