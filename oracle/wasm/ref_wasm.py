"""Driver for the reference's own compiled code: /root/reference/docs/bonnie-32.wasm run in wasm_interp.cpp.

TEST INFRASTRUCTURE (oracle side).  Only usable in the build container (the reference tree does not exist on
the GPU box); its job is to WRITE the fixtures under tests/golden/ref_wasm/ (make_ref_fixtures.py), which then
travel with the repo.
"""
import ctypes
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
WASM = '/root/reference/docs/bonnie-32.wasm'
LIB = os.path.join(HERE, '..', '_ref', 'libwasm_interp.so')


def build():
    src = os.path.join(HERE, 'wasm_interp.cpp')
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fno-fast-math', '-fPIC',
                               '-shared', '-o', LIB, src])


class WasmTrap(RuntimeError):
    pass


class RefWasm:
    def __init__(self, path=WASM):
        from wasmparse import Module
        build()
        self.mod = Module(path)
        L = self.lib = ctypes.CDLL(LIB)
        L.wi_load.restype = ctypes.c_void_p
        L.wi_load.argtypes = [ctypes.c_char_p]
        L.wi_call.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint32,
                              ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint32]
        L.wi_error.restype = ctypes.c_char_p
        L.wi_error.argtypes = [ctypes.c_void_p]
        L.wi_mem.restype = ctypes.c_void_p
        L.wi_mem.argtypes = [ctypes.c_void_p]
        L.wi_mem_size.restype = ctypes.c_uint64
        L.wi_mem_size.argtypes = [ctypes.c_void_p]
        L.wi_icount.restype = ctypes.c_uint64
        L.wi_icount.argtypes = [ctypes.c_void_p]
        L.wi_set_fuel.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
        L.wi_count_calls.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.wi_call_count.restype = ctypes.c_uint64
        L.wi_call_count.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.wi_watch.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
        L.wi_watch_count.restype = ctypes.c_uint64
        L.wi_watch_count.argtypes = [ctypes.c_void_p]
        L.wi_watch_log.restype = ctypes.POINTER(ctypes.c_uint64)
        L.wi_watch_log.argtypes = [ctypes.c_void_p]
        self.h = L.wi_load(path.encode())
        if not self.h:
            raise RuntimeError('cannot load ' + path)
        self._by_name = {}

    def func(self, key):
        """Index of the unique function whose (mangled) name contains `key`."""
        if key not in self._by_name:
            hits = self.mod.find(key)
            if len(hits) != 1:
                raise KeyError(f'{key}: {len(hits)} matches {hits[:5]}')
            self._by_name[key] = hits[0][0]
        return self._by_name[key]

    def call(self, key, *args):
        idx = key if isinstance(key, int) else self.func(key)
        params, results = self.mod.sig(idx)
        assert len(params) == len(args), (params, args)
        raw = (ctypes.c_uint64 * max(1, len(args)))()
        for i, (t, a) in enumerate(zip(params, args)):
            if t == 0x7F:
                raw[i] = int(a) & 0xFFFFFFFF
            elif t == 0x7E:
                raw[i] = int(a) & 0xFFFFFFFFFFFFFFFF
            elif t == 0x7D:
                raw[i] = struct.unpack('<I', struct.pack('<f', a))[0]
            else:
                raw[i] = struct.unpack('<Q', struct.pack('<d', a))[0]
        res = (ctypes.c_uint64 * max(1, len(results)))()
        rc = self.lib.wi_call(self.h, idx, raw, len(args), res, len(results))
        if rc:
            raise WasmTrap(self.lib.wi_error(self.h).decode())
        out = []
        for t, v in zip(results, res):
            if t == 0x7F:
                out.append(v & 0xFFFFFFFF)
            elif t == 0x7E:
                out.append(v)
            elif t == 0x7D:
                out.append(struct.unpack('<f', struct.pack('<I', v & 0xFFFFFFFF))[0])
            else:
                out.append(struct.unpack('<d', struct.pack('<Q', v))[0])
        return out[0] if len(out) == 1 else tuple(out)

    # -- linear memory ----------------------------------------------------------------------------
    def mem(self):
        n = self.lib.wi_mem_size(self.h)
        buf = (ctypes.c_uint8 * n).from_address(self.lib.wi_mem(self.h))
        return np.frombuffer(buf, dtype=np.uint8)

    def alloc(self, size, align=8):
        p = self.call('___rust_alloc', max(size, 1), align)
        if p == 0:
            raise MemoryError(size)
        return p

    def free(self, p, size, align=8):
        self.call('___rust_dealloc', p, max(size, 1), align)

    def write(self, addr, data):
        b = np.frombuffer(bytes(data) if not isinstance(data, np.ndarray) else data.tobytes(), dtype=np.uint8)
        self.mem()[addr:addr + len(b)] = b

    def read(self, addr, n):
        return bytes(self.mem()[addr:addr + n])

    def put(self, data, align=8):
        b = data.tobytes() if isinstance(data, np.ndarray) else bytes(data)
        p = self.alloc(len(b), align)
        self.write(p, b)
        return p

    @property
    def icount(self):
        return self.lib.wi_icount(self.h)

    # -- layout recovery ---------------------------------------------------------------------------
    def watch(self, lo, hi):
        self.lib.wi_watch(self.h, lo, hi)

    def watched(self):
        """Sorted unique (address, load opcode) pairs seen since watch()."""
        n = self.lib.wi_watch_count(self.h)
        p = self.lib.wi_watch_log(self.h)
        return sorted({(p[i] >> 16, p[i] & 0xFFFF) for i in range(n)})


# ---------------------------------------------------------------------------------------------------
# Addresses of macroquad statics read by the inlined `get_time()` prologue of render_mesh_15 / render_mesh
# (recovered from the function prologue, wasmdis.py render_mesh_15: `i64.load offset=1967488` etc.).
# get_time() asserts "called from the thread that owns the context" and "context exists" before reading
# context.start_time; outside the app both statics are unset, so the harness sets them by hand.
MQ_THREAD_ID = 1967488      # static THREAD_ID: Option<ThreadId> (0 = None)
MQ_TLS_INIT = 1967480       # thread-local "current thread id" lazy-init flag
MQ_TLS_THREAD_ID = 1967472  # thread-local current ThreadId
MQ_CONTEXT_TAG = 1963344    # static CONTEXT: Option<Context> discriminant (2 = None)


def enable_get_time(w):
    w.write(MQ_THREAD_ID, struct.pack('<Q', 1))
    w.write(MQ_TLS_INIT, b'\x01')
    w.write(MQ_TLS_THREAD_ID, struct.pack('<Q', 1))
    w.write(MQ_CONTEXT_TAG, struct.pack('<Q', 0))
