"""Minimal WebAssembly (MVP) binary reader: sections, types, imports, function names.

Test infrastructure (oracle side).  Used by ref_wasm.py to locate the reference's own compiled
functions inside /root/reference/docs/bonnie-32.wasm through the module's `name` section.
"""
import struct


def leb_u(d, p):
    r = 0
    s = 0
    while True:
        b = d[p]
        p += 1
        r |= (b & 0x7F) << s
        s += 7
        if not b & 0x80:
            return r, p


def leb_s(d, p, bits=64):
    r = 0
    s = 0
    while True:
        b = d[p]
        p += 1
        r |= (b & 0x7F) << s
        s += 7
        if not b & 0x80:
            if b & 0x40:
                r -= 1 << s
            return r, p


class Module:
    def __init__(self, path):
        self.data = d = open(path, 'rb').read()
        assert d[:8] == b'\0asm\1\0\0\0'
        self.sections = {}
        self.custom = {}
        p = 8
        while p < len(d):
            sid = d[p]
            p += 1
            sz, p = leb_u(d, p)
            if sid == 0:
                nl, q = leb_u(d, p)
                self.custom[d[q:q + nl].decode()] = (q + nl, p + sz)
            else:
                self.sections[sid] = (p, p + sz)
            p += sz
        self._types()
        self._imports()
        self._funcs()
        self._exports()
        self._names()

    def _types(self):
        d = self.data
        p, e = self.sections[1]
        n, p = leb_u(d, p)
        self.types = []
        for _ in range(n):
            assert d[p] == 0x60
            p += 1
            np_, p = leb_u(d, p)
            params = list(d[p:p + np_])
            p += np_
            nr, p = leb_u(d, p)
            res = list(d[p:p + nr])
            p += nr
            self.types.append((params, res))

    def _imports(self):
        d = self.data
        self.imports = []  # (module, name, kind, desc)
        if 2 not in self.sections:
            return
        p, e = self.sections[2]
        n, p = leb_u(d, p)
        for _ in range(n):
            l, p = leb_u(d, p)
            mod = d[p:p + l].decode()
            p += l
            l, p = leb_u(d, p)
            nm = d[p:p + l].decode()
            p += l
            kind = d[p]
            p += 1
            if kind == 0:
                t, p = leb_u(d, p)
                desc = t
            elif kind == 1:
                p += 1
                fl, p = leb_u(d, p)
                a, p = leb_u(d, p)
                if fl & 1:
                    b, p = leb_u(d, p)
                desc = None
            elif kind == 2:
                fl, p = leb_u(d, p)
                a, p = leb_u(d, p)
                b = None
                if fl & 1:
                    b, p = leb_u(d, p)
                desc = (a, b)
            else:
                p += 2
                desc = None
            self.imports.append((mod, nm, kind, desc))
        self.n_func_imports = sum(1 for i in self.imports if i[2] == 0)

    def _funcs(self):
        d = self.data
        p, e = self.sections[3]
        n, p = leb_u(d, p)
        self.func_types = [i[3] for i in self.imports if i[2] == 0]
        for _ in range(n):
            t, p = leb_u(d, p)
            self.func_types.append(t)

    def _exports(self):
        d = self.data
        self.exports = {}
        p, e = self.sections[7]
        n, p = leb_u(d, p)
        for _ in range(n):
            l, p = leb_u(d, p)
            nm = d[p:p + l].decode()
            p += l
            kind = d[p]
            p += 1
            idx, p = leb_u(d, p)
            self.exports[nm] = (kind, idx)

    def _names(self):
        d = self.data
        self.names = {}
        if 'name' not in self.custom:
            return
        p, e = self.custom['name']
        while p < e:
            sub = d[p]
            p += 1
            sz, p = leb_u(d, p)
            if sub == 1:
                q = p
                n, q = leb_u(d, q)
                for _ in range(n):
                    idx, q = leb_u(d, q)
                    l, q = leb_u(d, q)
                    self.names[idx] = d[q:q + l].decode(errors='replace')
                    q += l
            p += sz

    def find(self, substr):
        return [(i, n) for i, n in self.names.items() if substr in n]

    def sig(self, idx):
        return self.types[self.func_types[idx]]
