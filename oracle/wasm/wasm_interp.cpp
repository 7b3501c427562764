// wasm_interp.cpp — a small WebAssembly interpreter (MVP + sign-extension + saturating truncation +
// memory.copy/fill + multi-value block types), written for ONE job: executing the reference's own compiled
// rasterizer functions out of /root/reference/docs/bonnie-32.wasm (the authors' wasm32 release build, which
// keeps its `name` section) so that the C++ oracle in oracle/b32_oracle.cpp can be pinned against
// reference-EXECUTED results instead of a second reading of the source.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under bonnie-32_b200/ may link or load this; it is driven by
// oracle/wasm/ref_wasm.py (ctypes), which writes the fixtures under tests/golden/ref_wasm/.
//
// wasm f32/f64 arithmetic is strict IEEE-754 (round-to-nearest-even, no fusion, denormals kept): build with
// -O2 -ffp-contract=off -fno-fast-math.  f32.min/max/nearest/trunc follow the wasm spec (NaN-propagating min/max,
// -0 < +0).  NaN payload canonicalisation is not modelled (no observable effect on the rasterizer's integer
// outputs).
//
// Imports are stubs: every imported function returns zero(s) (`env.now` → 0.0), which is what the rasterizer's
// only host dependency (`get_time`, used for RasterTimings) needs.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct FuncType {
    std::vector<uint8_t> params, results;
};

enum : uint16_t {
    // pseudo ops produced by the decoder
    OP_FC_BASE = 0x100,  // 0xFC-prefixed ops are stored as 0x100 + sub
};

struct Instr {
    uint16_t op;
    uint16_t aux;   // block: nparams; br_table: unused
    uint32_t a;     // immediates: local/global/func index, label depth, memarg offset, else/end pc, nresults
    uint64_t imm;   // constants; block: (else_pc << 32) | end_pc
};

struct Func {
    uint32_t type = 0;
    bool imported = false;
    std::vector<uint8_t> local_types;  // beyond the parameters
    std::vector<Instr> code;
    std::vector<uint32_t> br_tables;   // flattened: [n, l0..ln-1, default]
    bool decoded = false;
    size_t body_begin = 0, body_end = 0;
};

struct Label {
    uint32_t cont_pc;
    uint32_t height;
    uint32_t arity;
    uint32_t is_loop;
};

struct Trap {
    const char* why;
};

struct Module {
    std::vector<uint8_t> bin;
    std::vector<FuncType> types;
    std::vector<Func> funcs;
    uint32_t n_import_funcs = 0;
    std::vector<uint64_t> globals;
    std::vector<uint32_t> table;  // function indices, 0xFFFFFFFF = null
    std::vector<uint8_t> mem;
    uint32_t mem_pages = 0, mem_max_pages = 65536;
    std::vector<uint64_t> stack;
    std::vector<Label> labels;
    uint64_t icount = 0;
    uint64_t fuel = 0;  // 0 = unlimited
    int depth = 0;
    std::string last_error;
    // execution trace hooks: count of calls per function (optional)
    std::vector<uint64_t> call_counts;
    bool count_calls = false;
    // layout recovery: every load whose address falls in [watch_lo, watch_hi) is logged as (addr, opcode)
    uint64_t watch_lo = 0, watch_hi = 0;
    std::vector<uint64_t> watch_log;
};

struct Reader {
    const uint8_t* d;
    size_t p, e;
    uint8_t u8() {
        if (p >= e) throw Trap{"read past end"};
        return d[p++];
    }
    uint64_t leb_u() {
        uint64_t r = 0;
        int s = 0;
        for (;;) {
            uint8_t b = u8();
            r |= (uint64_t)(b & 0x7F) << s;
            s += 7;
            if (!(b & 0x80)) return r;
        }
    }
    int64_t leb_s() {
        int64_t r = 0;
        int s = 0;
        for (;;) {
            uint8_t b = u8();
            r |= (int64_t)(b & 0x7F) << s;
            s += 7;
            if (!(b & 0x80)) {
                if ((b & 0x40) && s < 64) r |= -((int64_t)1 << s);
                return r;
            }
        }
    }
    uint32_t u32le() {
        uint32_t v;
        memcpy(&v, d + p, 4);
        p += 4;
        return v;
    }
    uint64_t u64le() {
        uint64_t v;
        memcpy(&v, d + p, 8);
        p += 8;
        return v;
    }
};

uint64_t const_expr(Module& m, Reader& r) {
    uint64_t v = 0;
    for (;;) {
        uint8_t op = r.u8();
        if (op == 0x0B) return v;
        if (op == 0x41) v = (uint32_t)(int32_t)r.leb_s();
        else if (op == 0x42) v = (uint64_t)r.leb_s();
        else if (op == 0x43) v = r.u32le();
        else if (op == 0x44) v = r.u64le();
        else if (op == 0x23) v = m.globals.at(r.leb_u());
        else if (op == 0xD2) v = r.leb_u();          // ref.func
        else if (op == 0xD0) { r.u8(); v = 0xFFFFFFFFu; }  // ref.null
        else throw Trap{"unsupported const expr"};
    }
}

void parse(Module& m) {
    Reader r{m.bin.data(), 0, m.bin.size()};
    if (m.bin.size() < 8 || memcmp(m.bin.data(), "\0asm\1\0\0\0", 8)) throw Trap{"not a wasm module"};
    r.p = 8;
    std::vector<uint32_t> func_type_idx;
    while (r.p < r.e) {
        uint8_t sid = r.u8();
        size_t sz = r.leb_u();
        size_t end = r.p + sz;
        Reader s{r.d, r.p, end};
        switch (sid) {
        case 1: {
            uint32_t n = s.leb_u();
            for (uint32_t i = 0; i < n; i++) {
                if (s.u8() != 0x60) throw Trap{"bad func type"};
                FuncType t;
                uint32_t np = s.leb_u();
                for (uint32_t k = 0; k < np; k++) t.params.push_back(s.u8());
                uint32_t nr = s.leb_u();
                for (uint32_t k = 0; k < nr; k++) t.results.push_back(s.u8());
                m.types.push_back(t);
            }
            break;
        }
        case 2: {
            uint32_t n = s.leb_u();
            for (uint32_t i = 0; i < n; i++) {
                uint32_t l = s.leb_u(); s.p += l;
                l = s.leb_u(); s.p += l;
                uint8_t kind = s.u8();
                if (kind == 0) {
                    Func f;
                    f.type = s.leb_u();
                    f.imported = true;
                    m.funcs.push_back(f);
                    m.n_import_funcs++;
                } else if (kind == 1) {
                    s.u8();
                    uint32_t fl = s.leb_u(); s.leb_u();
                    if (fl & 1) s.leb_u();
                } else if (kind == 2) {
                    uint32_t fl = s.leb_u();
                    m.mem_pages = s.leb_u();
                    if (fl & 1) m.mem_max_pages = s.leb_u();
                } else if (kind == 3) {
                    s.u8(); s.u8();
                    m.globals.push_back(0);
                }
            }
            break;
        }
        case 3: {
            uint32_t n = s.leb_u();
            for (uint32_t i = 0; i < n; i++) {
                Func f;
                f.type = s.leb_u();
                m.funcs.push_back(f);
            }
            break;
        }
        case 4: {
            uint32_t n = s.leb_u();
            for (uint32_t i = 0; i < n; i++) {
                s.u8();
                uint32_t fl = s.leb_u();
                uint32_t mn = s.leb_u();
                if (fl & 1) s.leb_u();
                if (i == 0) m.table.assign(mn, 0xFFFFFFFFu);
            }
            break;
        }
        case 5: {
            uint32_t n = s.leb_u();
            for (uint32_t i = 0; i < n; i++) {
                uint32_t fl = s.leb_u();
                m.mem_pages = s.leb_u();
                if (fl & 1) m.mem_max_pages = s.leb_u();
            }
            break;
        }
        case 6: {
            uint32_t n = s.leb_u();
            for (uint32_t i = 0; i < n; i++) {
                s.u8(); s.u8();
                m.globals.push_back(const_expr(m, s));
            }
            break;
        }
        case 9: {
            uint32_t n = s.leb_u();
            for (uint32_t i = 0; i < n; i++) {
                uint32_t fl = s.leb_u();
                if (fl == 0) {
                    uint32_t off = (uint32_t)const_expr(m, s);
                    uint32_t cnt = s.leb_u();
                    if (m.table.size() < off + cnt) m.table.resize(off + cnt, 0xFFFFFFFFu);
                    for (uint32_t k = 0; k < cnt; k++) m.table[off + k] = s.leb_u();
                } else if (fl == 2) {
                    s.leb_u();
                    uint32_t off = (uint32_t)const_expr(m, s);
                    s.u8();
                    uint32_t cnt = s.leb_u();
                    if (m.table.size() < off + cnt) m.table.resize(off + cnt, 0xFFFFFFFFu);
                    for (uint32_t k = 0; k < cnt; k++) m.table[off + k] = s.leb_u();
                } else if (fl == 1 || fl == 3) {  // passive / declarative: skip
                    s.u8();
                    uint32_t cnt = s.leb_u();
                    for (uint32_t k = 0; k < cnt; k++) s.leb_u();
                } else {
                    throw Trap{"unsupported element segment kind"};
                }
            }
            break;
        }
        case 10: {
            uint32_t n = s.leb_u();
            for (uint32_t i = 0; i < n; i++) {
                size_t bsz = s.leb_u();
                Func& f = m.funcs.at(m.n_import_funcs + i);
                f.body_begin = s.p;
                f.body_end = s.p + bsz;
                s.p += bsz;
            }
            break;
        }
        case 11: {
            m.mem.assign((size_t)m.mem_pages * 65536, 0);
            uint32_t n = s.leb_u();
            for (uint32_t i = 0; i < n; i++) {
                uint32_t fl = s.leb_u();
                if (fl == 1) {  // passive
                    uint32_t len = s.leb_u();
                    s.p += len;
                    continue;
                }
                if (fl == 2) s.leb_u();
                uint32_t off = (uint32_t)const_expr(m, s);
                uint32_t len = s.leb_u();
                if ((size_t)off + len > m.mem.size()) throw Trap{"data segment out of range"};
                memcpy(m.mem.data() + off, s.d + s.p, len);
                s.p += len;
            }
            break;
        }
        default:
            break;
        }
        r.p = end;
    }
    if (m.mem.empty()) m.mem.assign((size_t)m.mem_pages * 65536, 0);
}

// Decodes one function body into Instr[], resolving else/end positions of structured blocks.
void decode(Module& m, Func& f) {
    Reader r{m.bin.data(), f.body_begin, f.body_end};
    uint32_t ng = r.leb_u();
    for (uint32_t i = 0; i < ng; i++) {
        uint32_t c = r.leb_u();
        uint8_t t = r.u8();
        f.local_types.insert(f.local_types.end(), c, t);
    }
    std::vector<uint32_t> open;  // indices of block/loop/if instrs
    auto& code = f.code;
    while (r.p < r.e) {
        uint8_t op = r.u8();
        Instr in{op, 0, 0, 0};
        switch (op) {
        case 0x02: case 0x03: case 0x04: {
            int64_t bt = r.leb_s();
            uint32_t np = 0, nr = 0;
            if (bt >= 0) {
                np = m.types.at(bt).params.size();
                nr = m.types.at(bt).results.size();
            } else if ((bt & 0x7F) != 0x40) {
                nr = 1;
            }
            in.aux = np;
            in.a = nr;
            open.push_back(code.size());
            break;
        }
        case 0x05: {
            Instr& blk = code.at(open.back());
            blk.imm |= (uint64_t)code.size() << 32;  // else pc
            break;
        }
        case 0x0B: {
            if (!open.empty()) {
                Instr& blk = code.at(open.back());
                blk.imm |= (uint32_t)code.size();  // end pc
                open.pop_back();
                // an `else` needs to know its end: patch later through the block — store in else's a
                uint32_t else_pc = blk.imm >> 32;
                if (else_pc) code.at(else_pc).a = code.size();
            } else {
                in.op = 0x0F;  // function end == return
            }
            break;
        }
        case 0x0C: case 0x0D: in.a = r.leb_u(); break;
        case 0x0E: {
            uint32_t n = r.leb_u();
            in.a = f.br_tables.size();
            f.br_tables.push_back(n);
            for (uint32_t i = 0; i <= n; i++) f.br_tables.push_back(r.leb_u());
            break;
        }
        case 0x10: in.a = r.leb_u(); break;
        case 0x11: in.a = r.leb_u(); r.leb_u(); break;
        case 0x1C: { uint32_t n = r.leb_u(); r.p += n; in.op = 0x1B; break; }
        case 0x20: case 0x21: case 0x22: case 0x23: case 0x24: in.a = r.leb_u(); break;
        case 0x3F: case 0x40: r.u8(); break;
        case 0x41: in.imm = (uint32_t)(int32_t)r.leb_s(); break;
        case 0x42: in.imm = (uint64_t)r.leb_s(); break;
        case 0x43: in.imm = r.u32le(); break;
        case 0x44: in.imm = r.u64le(); break;
        case 0xFC: {
            uint32_t sub = r.leb_u();
            in.op = OP_FC_BASE + sub;
            if (sub == 10) { r.u8(); r.u8(); }
            else if (sub == 11) r.u8();
            else if (sub > 7) throw Trap{"unsupported 0xFC op"};
            break;
        }
        default:
            if (op >= 0x28 && op <= 0x3E) {
                r.leb_u();
                in.a = r.leb_u();
            } else if (!((op >= 0x45 && op <= 0xC4) || op == 0x00 || op == 0x01 || op == 0x0F || op == 0x1A ||
                         op == 0x1B)) {
                throw Trap{"unsupported opcode"};
            }
        }
        code.push_back(in);
    }
    f.decoded = true;
}

inline float f32_of(uint64_t v) { float f; uint32_t u = (uint32_t)v; memcpy(&f, &u, 4); return f; }
inline double f64_of(uint64_t v) { double f; memcpy(&f, &v, 8); return f; }
inline uint64_t of_f32(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline uint64_t of_f64(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }

template <class F> F wasm_min(F a, F b) {
    if (a != a || b != b) return std::numeric_limits<F>::quiet_NaN();
    if (a == 0 && b == 0) return std::signbit(a) ? a : b;
    return a < b ? a : b;
}
template <class F> F wasm_max(F a, F b) {
    if (a != a || b != b) return std::numeric_limits<F>::quiet_NaN();
    if (a == 0 && b == 0) return std::signbit(a) ? b : a;
    return a > b ? a : b;
}

template <class I, class F> I trunc_sat(F x) {
    if (x != x) return 0;
    const F lo = (F)std::numeric_limits<I>::min();
    // max+1 is exactly representable (a power of two)
    const F hi = (F)2 * (F)((uint64_t)1 << (sizeof(I) * 8 - 1 - (std::numeric_limits<I>::is_signed ? 1 : 0)));
    if (x <= lo) { if (x == lo || std::numeric_limits<I>::is_signed == false) return std::numeric_limits<I>::min(); return std::numeric_limits<I>::min(); }
    if (x >= hi) return std::numeric_limits<I>::max();
    return (I)x;
}
template <class I, class F> I trunc_trap(F x) {
    if (x != x) throw Trap{"invalid conversion to integer"};
    const F lo = (F)std::numeric_limits<I>::min();
    const F hi = (F)2 * (F)((uint64_t)1 << (sizeof(I) * 8 - 1 - (std::numeric_limits<I>::is_signed ? 1 : 0)));
    F t = std::trunc(x);
    if (t < lo || t >= hi) throw Trap{"integer overflow"};
    return (I)t;
}

void invoke(Module& m, uint32_t fidx, uint32_t sp_args);

#define MEMCHK(addr, n) \
    if ((uint64_t)(addr) + (n) > m.mem.size()) throw Trap{"out of bounds memory access"}

// Runs function `fidx`; its arguments are the top of m.stack starting at index `base`.  On return the results
// replace them (stack size = base + nresults).
void invoke(Module& m, uint32_t fidx, uint32_t base) {
    Func& f = m.funcs.at(fidx);
    const FuncType& ft = m.types[f.type];
    if (m.count_calls) m.call_counts[fidx]++;
    if (f.imported) {
        m.stack.resize(base);
        for (size_t i = 0; i < ft.results.size(); i++) m.stack.push_back(0);
        return;
    }
    if (!f.decoded) decode(m, f);
    if (++m.depth > 2000) throw Trap{"call stack exhausted"};
    const uint32_t nparams = ft.params.size();
    const uint32_t nlocals = nparams + f.local_types.size();
    m.stack.resize(base + nlocals, 0);
    for (uint32_t i = nparams; i < nlocals; i++) m.stack[base + i] = 0;
    const size_t label_base = m.labels.size();
    const Instr* code = f.code.data();
    const uint32_t ncode = f.code.size();
    uint32_t pc = 0;
    auto& st = m.stack;
    // operand stack lives above the locals
    auto push = [&](uint64_t v) { st.push_back(v); };
    auto pop = [&]() { uint64_t v = st.back(); st.pop_back(); return v; };
    auto do_branch = [&](uint32_t depth_) -> bool {  // returns true if the function returns
        size_t li = m.labels.size() - 1 - depth_;
        if (m.labels.size() < label_base + 1 + depth_) {
            return true;  // branch to the function's own label
        }
        Label L = m.labels[li];
        size_t top = st.size();
        for (uint32_t k = 0; k < L.arity; k++) st[L.height + k] = st[top - L.arity + k];
        st.resize(L.height + L.arity);
        if (L.is_loop) {
            m.labels.resize(li + 1);
        } else {
            m.labels.resize(li);
        }
        pc = L.cont_pc;
        return false;
    };
    for (;;) {
        if (pc >= ncode) break;
        const Instr& in = code[pc++];
        m.icount++;
        switch (in.op) {
        case 0x00: throw Trap{"unreachable executed"};
        case 0x01: break;
        case 0x02: {  // block
            m.labels.push_back(Label{(uint32_t)in.imm + 1, (uint32_t)(st.size() - in.aux), in.a, 0});
            break;
        }
        case 0x03: {  // loop
            m.labels.push_back(Label{pc, (uint32_t)(st.size() - in.aux), in.aux, 1});
            break;
        }
        case 0x04: {  // if
            uint32_t c = (uint32_t)pop();
            uint32_t end_pc = (uint32_t)in.imm, else_pc = in.imm >> 32;
            m.labels.push_back(Label{end_pc + 1, (uint32_t)(st.size() - in.aux), in.a, 0});
            if (!c) {
                if (else_pc) pc = else_pc + 1;
                else { pc = end_pc + 1; m.labels.pop_back(); }
            }
            break;
        }
        case 0x05: {  // else reached from the then-arm: skip to after end
            pc = in.a + 1;
            m.labels.pop_back();
            break;
        }
        case 0x0B: m.labels.pop_back(); break;
        case 0x0C:
            if (do_branch(in.a)) goto ret;
            break;
        case 0x0D:
            if ((uint32_t)pop()) { if (do_branch(in.a)) goto ret; }
            break;
        case 0x0E: {
            uint32_t i = (uint32_t)pop();
            const uint32_t* t = &f.br_tables[in.a];
            uint32_t n = t[0];
            uint32_t l = i < n ? t[1 + i] : t[1 + n];
            if (do_branch(l)) goto ret;
            break;
        }
        case 0x0F: goto ret;
        case 0x10: {
            const FuncType& ct = m.types[m.funcs[in.a].type];
            invoke(m, in.a, st.size() - ct.params.size());
            break;
        }
        case 0x11: {
            uint32_t ti = (uint32_t)pop();
            if (ti >= m.table.size() || m.table[ti] == 0xFFFFFFFFu) throw Trap{"undefined table element"};
            uint32_t callee = m.table[ti];
            const FuncType& want = m.types.at(in.a);
            const FuncType& have = m.types[m.funcs.at(callee).type];
            if (want.params != have.params || want.results != have.results) throw Trap{"indirect call type mismatch"};
            invoke(m, callee, st.size() - want.params.size());
            break;
        }
        case 0x1A: st.pop_back(); break;
        case 0x1B: {
            uint32_t c = (uint32_t)pop();
            uint64_t b = pop(), a = pop();
            push(c ? a : b);
            break;
        }
        case 0x20: push(st[base + in.a]); break;
        case 0x21: st[base + in.a] = st.back(); st.pop_back(); break;
        case 0x22: st[base + in.a] = st.back(); break;
        case 0x23: push(m.globals.at(in.a)); break;
        case 0x24: m.globals.at(in.a) = pop(); break;
#define LOAD(T, conv) { uint64_t ea = (uint64_t)(uint32_t)pop() + in.a; MEMCHK(ea, sizeof(T)); \
    if (ea >= m.watch_lo && ea < m.watch_hi && m.watch_log.size() < (1u << 22)) m.watch_log.push_back((ea << 16) | in.op); T v; memcpy(&v, &m.mem[ea], sizeof(T)); push(conv); break; }
        case 0x28: LOAD(uint32_t, (uint64_t)v)
        case 0x29: LOAD(uint64_t, v)
        case 0x2A: LOAD(uint32_t, (uint64_t)v)
        case 0x2B: LOAD(uint64_t, v)
        case 0x2C: LOAD(int8_t, (uint64_t)(uint32_t)(int32_t)v)
        case 0x2D: LOAD(uint8_t, (uint64_t)v)
        case 0x2E: LOAD(int16_t, (uint64_t)(uint32_t)(int32_t)v)
        case 0x2F: LOAD(uint16_t, (uint64_t)v)
        case 0x30: LOAD(int8_t, (uint64_t)(int64_t)v)
        case 0x31: LOAD(uint8_t, (uint64_t)v)
        case 0x32: LOAD(int16_t, (uint64_t)(int64_t)v)
        case 0x33: LOAD(uint16_t, (uint64_t)v)
        case 0x34: LOAD(int32_t, (uint64_t)(int64_t)v)
        case 0x35: LOAD(uint32_t, (uint64_t)v)
#define STORE(T) { T v = (T)pop(); uint64_t ea = (uint64_t)(uint32_t)pop() + in.a; MEMCHK(ea, sizeof(T)); memcpy(&m.mem[ea], &v, sizeof(T)); break; }
        case 0x36: STORE(uint32_t)
        case 0x37: STORE(uint64_t)
        case 0x38: STORE(uint32_t)
        case 0x39: STORE(uint64_t)
        case 0x3A: STORE(uint8_t)
        case 0x3B: STORE(uint16_t)
        case 0x3C: STORE(uint8_t)
        case 0x3D: STORE(uint16_t)
        case 0x3E: STORE(uint32_t)
        case 0x3F: push(m.mem.size() / 65536); break;
        case 0x40: {
            uint32_t n = (uint32_t)pop();
            uint64_t cur = m.mem.size() / 65536;
            if (cur + n > m.mem_max_pages || cur + n > 32768) { push(0xFFFFFFFFu); break; }
            m.mem.resize((cur + n) * 65536, 0);
            push(cur);
            break;
        }
        case 0x41: case 0x42: case 0x43: case 0x44: push(in.imm); break;
#define I32 (uint32_t)
#define S32 (int32_t)(uint32_t)
#define S64 (int64_t)
#define BIN(expr) { uint64_t b = pop(), a = pop(); (void)a; (void)b; push(expr); break; }
#define UN(expr) { uint64_t a = pop(); push(expr); break; }
        case 0x45: UN((uint64_t)(I32 a == 0))
        case 0x46: BIN((uint64_t)(I32 a == I32 b))
        case 0x47: BIN((uint64_t)(I32 a != I32 b))
        case 0x48: BIN((uint64_t)(S32 a < S32 b))
        case 0x49: BIN((uint64_t)(I32 a < I32 b))
        case 0x4A: BIN((uint64_t)(S32 a > S32 b))
        case 0x4B: BIN((uint64_t)(I32 a > I32 b))
        case 0x4C: BIN((uint64_t)(S32 a <= S32 b))
        case 0x4D: BIN((uint64_t)(I32 a <= I32 b))
        case 0x4E: BIN((uint64_t)(S32 a >= S32 b))
        case 0x4F: BIN((uint64_t)(I32 a >= I32 b))
        case 0x50: UN((uint64_t)(a == 0))
        case 0x51: BIN((uint64_t)(a == b))
        case 0x52: BIN((uint64_t)(a != b))
        case 0x53: BIN((uint64_t)(S64 a < S64 b))
        case 0x54: BIN((uint64_t)(a < b))
        case 0x55: BIN((uint64_t)(S64 a > S64 b))
        case 0x56: BIN((uint64_t)(a > b))
        case 0x57: BIN((uint64_t)(S64 a <= S64 b))
        case 0x58: BIN((uint64_t)(a <= b))
        case 0x59: BIN((uint64_t)(S64 a >= S64 b))
        case 0x5A: BIN((uint64_t)(a >= b))
        case 0x5B: BIN((uint64_t)(f32_of(a) == f32_of(b)))
        case 0x5C: BIN((uint64_t)(f32_of(a) != f32_of(b)))
        case 0x5D: BIN((uint64_t)(f32_of(a) < f32_of(b)))
        case 0x5E: BIN((uint64_t)(f32_of(a) > f32_of(b)))
        case 0x5F: BIN((uint64_t)(f32_of(a) <= f32_of(b)))
        case 0x60: BIN((uint64_t)(f32_of(a) >= f32_of(b)))
        case 0x61: BIN((uint64_t)(f64_of(a) == f64_of(b)))
        case 0x62: BIN((uint64_t)(f64_of(a) != f64_of(b)))
        case 0x63: BIN((uint64_t)(f64_of(a) < f64_of(b)))
        case 0x64: BIN((uint64_t)(f64_of(a) > f64_of(b)))
        case 0x65: BIN((uint64_t)(f64_of(a) <= f64_of(b)))
        case 0x66: BIN((uint64_t)(f64_of(a) >= f64_of(b)))
        case 0x67: UN((uint64_t)(I32 a ? __builtin_clz(I32 a) : 32))
        case 0x68: UN((uint64_t)(I32 a ? __builtin_ctz(I32 a) : 32))
        case 0x69: UN((uint64_t)__builtin_popcount(I32 a))
        case 0x6A: BIN((uint64_t)(I32(I32 a + I32 b)))
        case 0x6B: BIN((uint64_t)(I32(I32 a - I32 b)))
        case 0x6C: BIN((uint64_t)(I32(I32 a * I32 b)))
        case 0x6D: {
            int32_t b = S32 pop(), a = S32 pop();
            if (b == 0) throw Trap{"integer divide by zero"};
            if (a == INT32_MIN && b == -1) throw Trap{"integer overflow"};
            push((uint64_t)(uint32_t)(a / b));
            break;
        }
        case 0x6E: { uint32_t b = I32 pop(), a = I32 pop(); if (!b) throw Trap{"integer divide by zero"}; push(a / b); break; }
        case 0x6F: {
            int32_t b = S32 pop(), a = S32 pop();
            if (b == 0) throw Trap{"integer divide by zero"};
            push((uint64_t)(uint32_t)((a == INT32_MIN && b == -1) ? 0 : a % b));
            break;
        }
        case 0x70: { uint32_t b = I32 pop(), a = I32 pop(); if (!b) throw Trap{"integer divide by zero"}; push(a % b); break; }
        case 0x71: BIN((uint64_t)(I32 a & I32 b))
        case 0x72: BIN((uint64_t)(I32 a | I32 b))
        case 0x73: BIN((uint64_t)(I32 a ^ I32 b))
        case 0x74: BIN((uint64_t)(I32(I32 a << (b & 31))))
        case 0x75: BIN((uint64_t)(uint32_t)(S32 a >> (b & 31)))
        case 0x76: BIN((uint64_t)(I32 a >> (b & 31)))
        case 0x77: BIN((uint64_t)(I32((I32 a << (b & 31)) | (I32 a >> ((32 - (b & 31)) & 31)))))
        case 0x78: BIN((uint64_t)(I32((I32 a >> (b & 31)) | (I32 a << ((32 - (b & 31)) & 31)))))
        case 0x79: UN((uint64_t)(a ? __builtin_clzll(a) : 64))
        case 0x7A: UN((uint64_t)(a ? __builtin_ctzll(a) : 64))
        case 0x7B: UN((uint64_t)__builtin_popcountll(a))
        case 0x7C: BIN(a + b)
        case 0x7D: BIN(a - b)
        case 0x7E: BIN(a * b)
        case 0x7F: {
            int64_t b = S64 pop(), a = S64 pop();
            if (b == 0) throw Trap{"integer divide by zero"};
            if (a == INT64_MIN && b == -1) throw Trap{"integer overflow"};
            push((uint64_t)(a / b));
            break;
        }
        case 0x80: { uint64_t b = pop(), a = pop(); if (!b) throw Trap{"integer divide by zero"}; push(a / b); break; }
        case 0x81: {
            int64_t b = S64 pop(), a = S64 pop();
            if (b == 0) throw Trap{"integer divide by zero"};
            push((uint64_t)((a == INT64_MIN && b == -1) ? 0 : a % b));
            break;
        }
        case 0x82: { uint64_t b = pop(), a = pop(); if (!b) throw Trap{"integer divide by zero"}; push(a % b); break; }
        case 0x83: BIN(a & b)
        case 0x84: BIN(a | b)
        case 0x85: BIN(a ^ b)
        case 0x86: BIN(a << (b & 63))
        case 0x87: BIN((uint64_t)(S64 a >> (b & 63)))
        case 0x88: BIN(a >> (b & 63))
        case 0x89: BIN((a << (b & 63)) | (a >> ((64 - (b & 63)) & 63)))
        case 0x8A: BIN((a >> (b & 63)) | (a << ((64 - (b & 63)) & 63)))
        case 0x8B: UN(a & 0x7FFFFFFFu)
        case 0x8C: UN((a ^ 0x80000000u) & 0xFFFFFFFFu)
        case 0x8D: UN(of_f32(std::ceil(f32_of(a))))
        case 0x8E: UN(of_f32(std::floor(f32_of(a))))
        case 0x8F: UN(of_f32(std::trunc(f32_of(a))))
        case 0x90: UN(of_f32(std::nearbyint(f32_of(a))))
        case 0x91: UN(of_f32(std::sqrt(f32_of(a))))
        case 0x92: BIN(of_f32(f32_of(a) + f32_of(b)))
        case 0x93: BIN(of_f32(f32_of(a) - f32_of(b)))
        case 0x94: BIN(of_f32(f32_of(a) * f32_of(b)))
        case 0x95: BIN(of_f32(f32_of(a) / f32_of(b)))
        case 0x96: BIN(of_f32(wasm_min(f32_of(a), f32_of(b))))
        case 0x97: BIN(of_f32(wasm_max(f32_of(a), f32_of(b))))
        case 0x98: BIN((a & 0x7FFFFFFFu) | (b & 0x80000000u))
        case 0x99: UN(a & 0x7FFFFFFFFFFFFFFFull)
        case 0x9A: UN(a ^ 0x8000000000000000ull)
        case 0x9B: UN(of_f64(std::ceil(f64_of(a))))
        case 0x9C: UN(of_f64(std::floor(f64_of(a))))
        case 0x9D: UN(of_f64(std::trunc(f64_of(a))))
        case 0x9E: UN(of_f64(std::nearbyint(f64_of(a))))
        case 0x9F: UN(of_f64(std::sqrt(f64_of(a))))
        case 0xA0: BIN(of_f64(f64_of(a) + f64_of(b)))
        case 0xA1: BIN(of_f64(f64_of(a) - f64_of(b)))
        case 0xA2: BIN(of_f64(f64_of(a) * f64_of(b)))
        case 0xA3: BIN(of_f64(f64_of(a) / f64_of(b)))
        case 0xA4: BIN(of_f64(wasm_min(f64_of(a), f64_of(b))))
        case 0xA5: BIN(of_f64(wasm_max(f64_of(a), f64_of(b))))
        case 0xA6: BIN((a & 0x7FFFFFFFFFFFFFFFull) | (b & 0x8000000000000000ull))
        case 0xA7: UN(a & 0xFFFFFFFFu)
        case 0xA8: UN((uint64_t)(uint32_t)trunc_trap<int32_t>(f32_of(a)))
        case 0xA9: UN((uint64_t)trunc_trap<uint32_t>(f32_of(a)))
        case 0xAA: UN((uint64_t)(uint32_t)trunc_trap<int32_t>(f64_of(a)))
        case 0xAB: UN((uint64_t)trunc_trap<uint32_t>(f64_of(a)))
        case 0xAC: UN((uint64_t)(int64_t)(S32 a))
        case 0xAD: UN(a & 0xFFFFFFFFu)
        case 0xAE: UN((uint64_t)trunc_trap<int64_t>(f32_of(a)))
        case 0xAF: UN(trunc_trap<uint64_t>(f32_of(a)))
        case 0xB0: UN((uint64_t)trunc_trap<int64_t>(f64_of(a)))
        case 0xB1: UN(trunc_trap<uint64_t>(f64_of(a)))
        case 0xB2: UN(of_f32((float)(S32 a)))
        case 0xB3: UN(of_f32((float)(I32 a)))
        case 0xB4: UN(of_f32((float)(S64 a)))
        case 0xB5: UN(of_f32((float)a))
        case 0xB6: UN(of_f32((float)f64_of(a)))
        case 0xB7: UN(of_f64((double)(S32 a)))
        case 0xB8: UN(of_f64((double)(I32 a)))
        case 0xB9: UN(of_f64((double)(S64 a)))
        case 0xBA: UN(of_f64((double)a))
        case 0xBB: UN(of_f64((double)f32_of(a)))
        case 0xBC: case 0xBD: case 0xBE: case 0xBF: break;  // reinterpret: bits unchanged
        case 0xC0: UN((uint64_t)(uint32_t)(int32_t)(int8_t)a)
        case 0xC1: UN((uint64_t)(uint32_t)(int32_t)(int16_t)a)
        case 0xC2: UN((uint64_t)(int64_t)(int8_t)a)
        case 0xC3: UN((uint64_t)(int64_t)(int16_t)a)
        case 0xC4: UN((uint64_t)(int64_t)(int32_t)a)
        case OP_FC_BASE + 0: UN((uint64_t)(uint32_t)trunc_sat<int32_t>(f32_of(a)))
        case OP_FC_BASE + 1: UN((uint64_t)trunc_sat<uint32_t>(f32_of(a)))
        case OP_FC_BASE + 2: UN((uint64_t)(uint32_t)trunc_sat<int32_t>(f64_of(a)))
        case OP_FC_BASE + 3: UN((uint64_t)trunc_sat<uint32_t>(f64_of(a)))
        case OP_FC_BASE + 4: UN((uint64_t)trunc_sat<int64_t>(f32_of(a)))
        case OP_FC_BASE + 5: UN(trunc_sat<uint64_t>(f32_of(a)))
        case OP_FC_BASE + 6: UN((uint64_t)trunc_sat<int64_t>(f64_of(a)))
        case OP_FC_BASE + 7: UN(trunc_sat<uint64_t>(f64_of(a)))
        case OP_FC_BASE + 10: {
            uint32_t n = I32 pop(), s = I32 pop(), d = I32 pop();
            MEMCHK(s, n); MEMCHK(d, n);
            memmove(&m.mem[d], &m.mem[s], n);
            break;
        }
        case OP_FC_BASE + 11: {
            uint32_t n = I32 pop(), v = I32 pop(), d = I32 pop();
            MEMCHK(d, n);
            memset(&m.mem[d], (int)(v & 0xFF), n);
            break;
        }
        default: throw Trap{"unimplemented opcode at run time"};
        }
        if (m.fuel && m.icount > m.fuel) throw Trap{"out of fuel"};
    }
ret:
    {
        uint32_t nres = ft.results.size();
        size_t top = st.size();
        for (uint32_t k = 0; k < nres; k++) st[base + k] = st[top - nres + k];
        st.resize(base + nres);
        m.labels.resize(label_base);
        m.depth--;
    }
}

}  // namespace

extern "C" {

void* wi_load(const char* path) {
    FILE* fp = fopen(path, "rb");
    if (!fp) return nullptr;
    Module* m = new Module();
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    m->bin.resize(n);
    if (fread(m->bin.data(), 1, n, fp) != (size_t)n) { fclose(fp); delete m; return nullptr; }
    fclose(fp);
    try {
        parse(*m);
    } catch (Trap& t) {
        fprintf(stderr, "wi_load: %s\n", t.why);
        delete m;
        return nullptr;
    }
    m->stack.reserve(1 << 20);
    m->call_counts.assign(m->funcs.size(), 0);
    return m;
}

void wi_free(void* h) { delete (Module*)h; }

// Calls function `fidx` with nargs raw 64-bit argument slots (i32/f32 in the low 32 bits).  Returns 0 on
// success, 1 on a trap (message through wi_error).
int wi_call(void* h, uint32_t fidx, const uint64_t* args, uint32_t nargs, uint64_t* results, uint32_t nresults) {
    Module& m = *(Module*)h;
    try {
        if (fidx >= m.funcs.size()) throw Trap{"no such function"};
        const FuncType& ft = m.types[m.funcs[fidx].type];
        if (ft.params.size() != nargs || ft.results.size() != nresults) throw Trap{"signature mismatch"};
        m.stack.clear();
        m.labels.clear();
        m.depth = 0;
        for (uint32_t i = 0; i < nargs; i++) m.stack.push_back(args[i]);
        invoke(m, fidx, 0);
        for (uint32_t i = 0; i < nresults; i++) results[i] = m.stack[i];
        return 0;
    } catch (Trap& t) {
        m.last_error = t.why;
        return 1;
    } catch (std::exception& e) {
        m.last_error = e.what();
        return 1;
    }
}

const char* wi_error(void* h) { return ((Module*)h)->last_error.c_str(); }
uint8_t* wi_mem(void* h) { return ((Module*)h)->mem.data(); }
uint64_t wi_mem_size(void* h) { return ((Module*)h)->mem.size(); }
uint64_t wi_icount(void* h) { return ((Module*)h)->icount; }
void wi_set_fuel(void* h, uint64_t fuel) { ((Module*)h)->fuel = fuel ? ((Module*)h)->icount + fuel : 0; }
uint64_t wi_global_get(void* h, uint32_t i) { return ((Module*)h)->globals.at(i); }
void wi_global_set(void* h, uint32_t i, uint64_t v) { ((Module*)h)->globals.at(i) = v; }
void wi_count_calls(void* h, int on) {
    Module& m = *(Module*)h;
    m.count_calls = on != 0;
    if (on) std::fill(m.call_counts.begin(), m.call_counts.end(), 0);
}
uint64_t wi_call_count(void* h, uint32_t fidx) { return ((Module*)h)->call_counts.at(fidx); }
void wi_watch(void* h, uint64_t lo, uint64_t hi) {
    Module& m = *(Module*)h;
    m.watch_lo = lo;
    m.watch_hi = hi;
    m.watch_log.clear();
}
uint64_t wi_watch_count(void* h) { return ((Module*)h)->watch_log.size(); }
const uint64_t* wi_watch_log(void* h) { return ((Module*)h)->watch_log.data(); }
uint32_t wi_num_funcs(void* h) { return ((Module*)h)->funcs.size(); }

}  // extern "C"
