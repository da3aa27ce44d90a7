"""Generates tests/golden/vt_pyc_spec.json from the reference's orphaned CPython-3.7 bytecode
(/root/reference/nn_encoder_arch/__pycache__/vision_transformer.cpython-37.pyc).

The container only has CPython 3.12, whose `marshal` cannot rebuild 3.7 code objects, so the marshal
stream is parsed by hand and the 3.7 wordcode is walked just far enough to recover, for
every function / method: its argument names, its constant defaults (where the defaults are a constant
tuple), the global / attribute names it references and its numeric + string constants. That is the
only machine-checkable ground truth the reference offers for this path; tests/test_oracle_structure.py
pins the oracle and the drop-in modules to it. Run in the build container only (needs /root/reference).
"""
import json
import os
import struct
import sys

PYC = "/root/reference/nn_encoder_arch/__pycache__/vision_transformer.cpython-37.pyc"
FLAG_REF = 0x80


class Code:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Reader:
    def __init__(self, data):
        self.d, self.p, self.refs = data, 0, []

    def byte(self):
        b = self.d[self.p]
        self.p += 1
        return b

    def i32(self):
        v = struct.unpack_from("<i", self.d, self.p)[0]
        self.p += 4
        return v

    def read(self, n):
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def obj(self):
        code = self.byte()
        flag = code & FLAG_REF
        t = chr(code & ~FLAG_REF)
        idx = None

        def reserve():
            nonlocal idx
            if flag:
                idx = len(self.refs)
                self.refs.append(None)

        def done(v):
            if flag:
                if idx is None:
                    self.refs.append(v)
                else:
                    self.refs[idx] = v
            return v

        if t == "N":
            return None
        if t == "T":
            return True
        if t == "F":
            return False
        if t == ".":
            return Ellipsis
        if t == "i":
            return done(self.i32())
        if t == "l":
            n = self.i32()
            digits = [struct.unpack_from("<H", self.read(2))[0] for _ in range(abs(n))]
            v = sum(dg << (15 * i) for i, dg in enumerate(digits))
            return done(-v if n < 0 else v)
        if t == "g":
            return done(struct.unpack("<d", self.read(8))[0])
        if t == "y":
            return done(complex(*struct.unpack("<dd", self.read(16))))
        if t in "su":
            n = self.i32()
            b = self.read(n)
            return done(b if t == "s" else b.decode("utf8", "surrogatepass"))
        if t == "t":
            n = self.i32()
            return done(self.read(n).decode("utf8"))
        if t in "aA":
            n = self.i32()
            return done(self.read(n).decode("latin1"))
        if t in "zZ":
            n = self.byte()
            return done(self.read(n).decode("latin1"))
        if t in "()":
            n = self.byte() if t == ")" else self.i32()
            reserve()
            return done(tuple(self.obj() for _ in range(n)))
        if t == "[":
            n = self.i32()
            reserve()
            return done([self.obj() for _ in range(n)])
        if t in "<>":
            n = self.i32()
            reserve()
            return done(frozenset(self.obj() for _ in range(n)))
        if t == "{":
            reserve()
            out = {}
            while True:
                k = self.obj_or_null()
                if k is _NULL:
                    break
                out[k] = self.obj()
            return done(out)
        if t == "r":
            return self.refs[self.i32()]
        if t == "c":
            reserve()
            argcount, kwonly, nlocals, stacksize, flags = (self.i32() for _ in range(5))
            co = Code(argcount=argcount, kwonlyargcount=kwonly, nlocals=nlocals, flags=flags,
                      code=self.obj(), consts=self.obj(), names=self.obj(), varnames=self.obj(),
                      freevars=self.obj(), cellvars=self.obj(), filename=self.obj(), name=self.obj(),
                      firstlineno=self.i32(), lnotab=self.obj())
            return done(co)
        raise ValueError(f"unknown marshal type {t!r} at {self.p - 1}")

    def obj_or_null(self):
        if self.d[self.p] == ord("0"):
            self.p += 1
            return _NULL
        return self.obj()


_NULL = object()
LOAD_CONST, MAKE_FUNCTION, EXTENDED_ARG, LOAD_BUILD_CLASS = 100, 132, 144, 71


def walk(co, prefix, out):
    """Records this code object's functions; recurses into nested code constants."""
    # pass 1: find (defaults tuple, code, qualname) triples via the LOAD_CONST .. MAKE_FUNCTION pattern
    ops = []
    ext = 0
    bc = co.code
    for i in range(0, len(bc), 2):
        op, arg = bc[i], bc[i + 1] | ext
        ext = (arg << 8) if op == EXTENDED_ARG else 0
        if op != EXTENDED_ARG:
            ops.append((op, arg))
    defaults_of = {}
    for k, (op, arg) in enumerate(ops):
        if op == MAKE_FUNCTION and k >= 2 and ops[k - 1][0] == LOAD_CONST and ops[k - 2][0] == LOAD_CONST:
            code_c = co.consts[ops[k - 2][1]]
            if not isinstance(code_c, Code) or not (arg & 0x01):
                continue
            # stack below the code object, top first: [closure tuple] [annotations map] [kwdefaults] defaults
            j = k - 3
            if arg & 0x08:  # LOAD_CLOSURE ... BUILD_TUPLE n
                n = ops[j][1]
                j -= 1 + n
            if arg & 0x04:  # values..., LOAD_CONST keys, BUILD_CONST_KEY_MAP n
                n = ops[j][1]
                j -= 2 + n
            if j >= 0 and ops[j][0] == LOAD_CONST and isinstance(co.consts[ops[j][1]], tuple):
                defaults_of[id(code_c)] = co.consts[ops[j][1]]
    for c in co.consts:
        if isinstance(c, Code):
            qual = f"{prefix}{c.name}"
            is_class_body = "__qualname__" in c.names and "__module__" in c.names
            if not is_class_body:
                consts = [x for x in c.consts if isinstance(x, (int, float, str)) and not isinstance(x, bool)]
                entry = {"args": list(c.varnames[:c.argcount + c.kwonlyargcount]),
                         "varargs": bool(c.flags & 0x04), "varkw": bool(c.flags & 0x08),
                         "names": sorted(set(c.names)),
                         "num_consts": sorted({x for x in consts if isinstance(x, (int, float))}),
                         "str_consts": sorted({x for x in consts if isinstance(x, str) and len(x) < 40}),
                         "firstlineno": c.firstlineno}
                if id(c) in defaults_of:
                    entry["const_defaults"] = [x if isinstance(x, (int, float, str, bool, type(None))) else repr(x)
                                               for x in defaults_of[id(c)]]
                out[qual] = entry
            walk(c, qual + ".", out)


def main():
    data = open(PYC, "rb").read()
    magic = struct.unpack_from("<H", data, 0)[0]
    assert magic == 3394, magic
    top = Reader(data[16:]).obj()
    out = {}
    walk(top, "", out)
    spec = {"source": PYC, "magic": magic, "module_names": sorted(set(top.names)), "functions": out}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vt_pyc_spec.json")
    with open(dst, "w") as f:
        json.dump(spec, f, indent=1, sort_keys=True)
    print(f"wrote {dst}: {len(out)} functions")


if __name__ == "__main__":
    sys.exit(main())
