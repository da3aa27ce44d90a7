"""A small CPython-3.7 bytecode interpreter (test infrastructure only).

The reference ships its encoder / projection head only as CPython-3.7 bytecode
(/root/reference/nn_encoder_arch/__pycache__/vision_transformer.cpython-37.pyc); this container has CPython 3.12,
which can neither unmarshal nor execute it. This module executes that bytecode with 3.7 semantics on top of the
running interpreter: the marshal stream is parsed by ``make_vt_pyc_spec.Reader`` and the wordcode is run by the
stack machine below, so the reference's OWN code (its classes become real ``nn.Module`` subclasses, its functions real
callables) produces the golden vectors that pin the oracle (``make_vt_goldens.py``).

Covered: the 56 opcodes the file uses (straight-line code, loops, list comprehensions, closures, classes with
zero-argument ``super()``, ``with`` blocks (left by fall-through, ``return`` or a propagating exception), imports, calls with keyword / star
arguments). Exceptions raised by executed code propagate as ordinary Python exceptions; ``try/except`` inside the
bytecode is not supported (the file has none).
"""
from __future__ import annotations

import builtins
import operator
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_vt_pyc_spec import Code, Reader  # noqa: E402

# CPython 3.7 opcode numbers (Lib/opcode.py of 3.7)
POP_TOP, ROT_TWO, ROT_THREE, DUP_TOP, DUP_TOP_TWO, NOP = 1, 2, 3, 4, 5, 9
UNARY_POSITIVE, UNARY_NEGATIVE, UNARY_NOT, UNARY_INVERT = 10, 11, 12, 15
BINARY = {16: operator.matmul, 19: operator.pow, 20: operator.mul, 22: operator.mod, 23: operator.add,
          24: operator.sub, 25: operator.getitem, 26: operator.floordiv, 27: operator.truediv,
          62: operator.lshift, 63: operator.rshift, 64: operator.and_, 65: operator.xor, 66: operator.or_}
INPLACE = {17: operator.imatmul, 28: operator.ifloordiv, 29: operator.itruediv, 55: operator.iadd,
           56: operator.isub, 57: operator.imul, 59: operator.imod, 67: operator.ipow, 75: operator.ilshift,
           76: operator.irshift, 77: operator.iand, 78: operator.ixor, 79: operator.ior}
STORE_SUBSCR, DELETE_SUBSCR, GET_ITER, LOAD_BUILD_CLASS = 60, 61, 68, 71
WITH_CLEANUP_START, WITH_CLEANUP_FINISH, RETURN_VALUE, POP_BLOCK, END_FINALLY = 81, 82, 83, 87, 88
STORE_NAME, UNPACK_SEQUENCE, FOR_ITER, STORE_ATTR, STORE_GLOBAL = 90, 92, 93, 95, 97
LOAD_CONST, LOAD_NAME, BUILD_TUPLE, BUILD_LIST, BUILD_SET, BUILD_MAP, LOAD_ATTR, COMPARE_OP = 100, 101, 102, 103, 104, 105, 106, 107
IMPORT_NAME, IMPORT_FROM, JUMP_FORWARD, JUMP_IF_FALSE_OR_POP, JUMP_IF_TRUE_OR_POP, JUMP_ABSOLUTE = 108, 109, 110, 111, 112, 113
POP_JUMP_IF_FALSE, POP_JUMP_IF_TRUE, LOAD_GLOBAL, SETUP_LOOP = 114, 115, 116, 120
LOAD_FAST, STORE_FAST, DELETE_FAST, RAISE_VARARGS, CALL_FUNCTION, MAKE_FUNCTION, BUILD_SLICE = 124, 125, 126, 130, 131, 132, 133
LOAD_CLOSURE, LOAD_DEREF, STORE_DEREF, CALL_FUNCTION_KW, CALL_FUNCTION_EX, SETUP_WITH, EXTENDED_ARG = 135, 136, 137, 141, 142, 143, 144
LIST_APPEND, SET_ADD, MAP_ADD, BUILD_MAP_UNPACK_WITH_CALL, BUILD_TUPLE_UNPACK, BUILD_CONST_KEY_MAP = 145, 146, 147, 151, 152, 156
BUILD_STRING, BUILD_TUPLE_UNPACK_WITH_CALL, LOAD_METHOD, CALL_METHOD, FORMAT_VALUE = 157, 158, 160, 161, 155

CMP = [operator.lt, operator.le, operator.eq, operator.ne, operator.gt, operator.ge,
       lambda a, b: a in b, lambda a, b: a not in b, operator.is_, operator.is_not]


class Cell:
    __slots__ = ("value",)

    def __init__(self, value=None):
        self.value = value


_NULL = object()   # LOAD_METHOD's "not a method" marker


def make_function(code: Code, globs: dict, defaults=(), kwdefaults=None, closure=()):
    """A real Python function (so that it binds as a method) whose body runs the 3.7 code object."""
    nargs, nkw = code.argcount, code.kwonlyargcount
    names = code.varnames
    has_varargs, has_varkw = bool(code.flags & 0x04), bool(code.flags & 0x08)

    def call(*args, **kwargs):
        fast = {}
        if len(args) > nargs and not has_varargs:
            raise TypeError(f"{code.name}() takes {nargs} positional arguments but {len(args)} were given")
        for i, a in enumerate(args[:nargs]):
            fast[names[i]] = a
        slot = nargs + nkw
        if has_varargs:
            fast[names[slot]] = tuple(args[nargs:])
            slot += 1
        extra = {}
        for k, v in kwargs.items():
            if k in names[:nargs + nkw]:
                if k in fast:
                    raise TypeError(f"{code.name}() got multiple values for argument {k!r}")
                fast[k] = v
            elif has_varkw:
                extra[k] = v
            else:
                raise TypeError(f"{code.name}() got an unexpected keyword argument {k!r}")
        if has_varkw:
            fast[names[slot]] = extra
        first_default = nargs - len(defaults)
        for i in range(nargs):
            if names[i] not in fast:
                if i < first_default:
                    raise TypeError(f"{code.name}() missing required argument {names[i]!r}")
                fast[names[i]] = defaults[i - first_default]
        for i in range(nargs, nargs + nkw):
            if names[i] not in fast:
                if kwdefaults is None or names[i] not in kwdefaults:
                    raise TypeError(f"{code.name}() missing keyword-only argument {names[i]!r}")
                fast[names[i]] = kwdefaults[names[i]]
        return run_frame(code, globs, fast, closure)

    call.__name__ = code.name
    call.__qualname__ = code.name
    call.__py37_code__ = code
    return call


def build_class(body_fn, name, *bases, **kwds):
    """__build_class__ for interpreted class bodies: run the body into a namespace, create the type, and fill the
    ``__class__`` cell that zero-argument super() reads."""
    code = body_fn.__py37_code__
    meta = kwds.pop("metaclass", None) or (type(bases[0]) if bases else type)
    ns = meta.__prepare__(name, bases, **kwds) if hasattr(meta, "__prepare__") else {}
    cells = tuple(Cell() for _ in code.cellvars)
    run_frame(code, body_fn.__py37_globals__, {}, (), namespace=ns, own_cells=cells)
    classcell = ns.pop("__classcell__", None)
    cls = meta(name, bases, dict(ns), **kwds)
    if classcell is not None:
        classcell.value = cls
    return cls


def run_frame(code: Code, globs: dict, fast: dict, closure=(), namespace=None, own_cells=None):
    """Runs one code object. ``blocks`` mirrors ceval's block stack just far enough to leave ``with`` blocks
    correctly on ``return`` and on a propagating exception (3.7 unwinds them through WHY_RETURN / WHY_EXCEPTION)."""
    blocks = []
    try:
        return _run(code, globs, fast, closure, namespace, own_cells, blocks)
    except BaseException as e:
        for b in reversed(blocks):
            if b is not None:
                b(type(e), e, e.__traceback__)
        raise


def _run(code, globs, fast, closure, namespace, own_cells, blocks):
    bc = code.code
    consts, names, varnames = code.consts, code.names, code.varnames
    cells = list(own_cells) if own_cells is not None else [Cell() for _ in code.cellvars]
    # arguments that are also cell variables live in the cell
    for i, cv in enumerate(code.cellvars):
        if cv in fast:
            cells[i].value = fast[cv]
    deref = cells + list(closure)
    deref_names = list(code.cellvars) + list(code.freevars)
    stack = []
    push, pop = stack.append, stack.pop
    pc, ext = 0, 0
    n = len(bc)
    while pc < n:
        op, arg = bc[pc], bc[pc + 1] | ext
        pc += 2
        if op == EXTENDED_ARG:
            ext = arg << 8
            continue
        ext = 0
        if op == LOAD_FAST:
            try:
                push(fast[varnames[arg]])
            except KeyError:
                raise UnboundLocalError(varnames[arg]) from None
        elif op == LOAD_CONST:
            push(consts[arg])
        elif op == LOAD_GLOBAL:
            nm = names[arg]
            if nm in globs:
                push(globs[nm])
            else:
                push(getattr(builtins, nm))
        elif op == LOAD_ATTR:
            stack[-1] = getattr(stack[-1], names[arg])
        elif op == LOAD_METHOD:
            obj = pop()
            push(_NULL)
            push(getattr(obj, names[arg]))
        elif op == CALL_METHOD:
            args = [pop() for _ in range(arg)][::-1]
            fn = pop()
            pop()  # the _NULL marker
            push(fn(*args))
        elif op == STORE_FAST:
            fast[varnames[arg]] = pop()
        elif op == CALL_FUNCTION:
            args = [pop() for _ in range(arg)][::-1]
            fn = pop()
            if fn is builtins.super and not args:
                # zero-argument super(): the class comes from the frame's __class__ cell, the instance is argument 0
                cls = deref[deref_names.index("__class__")].value
                push(super(cls, fast[varnames[0]]))
            else:
                push(fn(*args))
        elif op == CALL_FUNCTION_KW:
            kwnames = pop()
            vals = [pop() for _ in range(arg)][::-1]
            fn = pop()
            nk = len(kwnames)
            push(fn(*vals[:arg - nk], **dict(zip(kwnames, vals[arg - nk:]))))
        elif op == CALL_FUNCTION_EX:
            kw = pop() if arg & 1 else {}
            pos = pop()
            fn = pop()
            push(fn(*pos, **kw))
        elif op == POP_TOP:
            pop()
        elif op == RETURN_VALUE:
            while blocks:
                b = blocks.pop()
                if b is not None:
                    b(None, None, None)
            return pop()
        elif op == STORE_ATTR:
            obj = pop()
            setattr(obj, names[arg], pop())
        elif op in BINARY:
            b = pop()
            stack[-1] = BINARY[op](stack[-1], b)
        elif op in INPLACE:
            b = pop()
            stack[-1] = INPLACE[op](stack[-1], b)
        elif op == COMPARE_OP:
            b = pop()
            stack[-1] = CMP[arg](stack[-1], b)
        elif op == POP_JUMP_IF_FALSE:
            if not pop():
                pc = arg
        elif op == POP_JUMP_IF_TRUE:
            if pop():
                pc = arg
        elif op == JUMP_IF_TRUE_OR_POP:
            if stack[-1]:
                pc = arg
            else:
                pop()
        elif op == JUMP_IF_FALSE_OR_POP:
            if not stack[-1]:
                pc = arg
            else:
                pop()
        elif op == JUMP_FORWARD:
            pc += arg
        elif op == JUMP_ABSOLUTE:
            pc = arg
        elif op == BUILD_TUPLE:
            vals = tuple(stack[len(stack) - arg:]) if arg else ()
            del stack[len(stack) - arg:]
            push(vals)
        elif op == BUILD_LIST:
            vals = list(stack[len(stack) - arg:]) if arg else []
            del stack[len(stack) - arg:]
            push(vals)
        elif op == BUILD_SLICE:
            step = pop() if arg == 3 else None
            stop = pop()
            stack[-1] = slice(stack[-1], stop, step)
        elif op == BUILD_CONST_KEY_MAP:
            keys = pop()
            vals = stack[len(stack) - arg:]
            del stack[len(stack) - arg:]
            push(dict(zip(keys, vals)))
        elif op == BUILD_MAP:
            items = stack[len(stack) - 2 * arg:]
            del stack[len(stack) - 2 * arg:]
            push({items[2 * i]: items[2 * i + 1] for i in range(arg)})
        elif op == BUILD_MAP_UNPACK_WITH_CALL:
            maps = stack[len(stack) - arg:]
            del stack[len(stack) - arg:]
            merged = {}
            for m in maps:
                for k in m:
                    if k in merged:
                        raise TypeError(f"got multiple values for keyword argument {k!r}")
                    merged[k] = m[k]
            push(merged)
        elif op in (BUILD_TUPLE_UNPACK, BUILD_TUPLE_UNPACK_WITH_CALL):
            parts = stack[len(stack) - arg:]
            del stack[len(stack) - arg:]
            push(tuple(x for p in parts for x in p))
        elif op == UNPACK_SEQUENCE:
            seq = list(pop())
            if len(seq) != arg:
                raise ValueError(f"expected {arg} values to unpack, got {len(seq)}")
            stack.extend(reversed(seq))
        elif op == GET_ITER:
            stack[-1] = iter(stack[-1])
        elif op == FOR_ITER:
            try:
                push(next(stack[-1]))
            except StopIteration:
                pop()
                pc += arg
        elif op == LIST_APPEND:
            v = pop()
            stack[-arg].append(v)
        elif op == SETUP_LOOP:
            blocks.append(None)
        elif op == POP_BLOCK:
            blocks.pop()
        elif op == NOP:
            pass
        elif op == DUP_TOP:
            push(stack[-1])
        elif op == DUP_TOP_TWO:
            stack.extend(stack[-2:])
        elif op == ROT_TWO:
            stack[-1], stack[-2] = stack[-2], stack[-1]
        elif op == ROT_THREE:
            stack[-1], stack[-2], stack[-3] = stack[-2], stack[-3], stack[-1]
        elif op == UNARY_NEGATIVE:
            stack[-1] = -stack[-1]
        elif op == UNARY_NOT:
            stack[-1] = not stack[-1]
        elif op == UNARY_POSITIVE:
            stack[-1] = +stack[-1]
        elif op == UNARY_INVERT:
            stack[-1] = ~stack[-1]
        elif op == STORE_SUBSCR:
            key = pop()
            obj = pop()
            obj[key] = pop()
        elif op == LOAD_CLOSURE:
            push(deref[arg])
        elif op == LOAD_DEREF:
            push(deref[arg].value)
        elif op == STORE_DEREF:
            deref[arg].value = pop()
        elif op == MAKE_FUNCTION:
            pop()  # qualified name
            fcode = pop()
            fclosure = pop() if arg & 0x08 else ()
            if arg & 0x04:
                pop()  # annotations
            kwdefaults = pop() if arg & 0x02 else None
            defaults = pop() if arg & 0x01 else ()
            fn = make_function(fcode, globs, defaults, kwdefaults, fclosure)
            fn.__py37_globals__ = globs
            push(fn)
        elif op == LOAD_BUILD_CLASS:
            push(build_class)
        elif op == STORE_NAME:
            (namespace if namespace is not None else globs)[names[arg]] = pop()
        elif op == LOAD_NAME:
            nm = names[arg]
            if namespace is not None and nm in namespace:
                push(namespace[nm])
            elif nm in globs:
                push(globs[nm])
            else:
                push(getattr(builtins, nm))
        elif op == STORE_GLOBAL:
            globs[names[arg]] = pop()
        elif op == IMPORT_NAME:
            fromlist = pop()
            level = pop()
            push(__import__(names[arg], globs, None, fromlist, level))
        elif op == IMPORT_FROM:
            push(getattr(stack[-1], names[arg]))
        elif op == RAISE_VARARGS:
            if arg == 1:
                exc = pop()
                raise exc
            if arg == 2:
                cause = pop()
                exc = pop()
                raise exc from cause
            raise RuntimeError("bare raise outside a handler")
        elif op == SETUP_WITH:
            mgr = pop()
            exit_fn = type(mgr).__exit__.__get__(mgr)
            push(exit_fn)
            push(type(mgr).__enter__(mgr))
            blocks.append(exit_fn)
        elif op == WITH_CLEANUP_START:
            exc = pop()          # None on the non-exceptional path (the only one supported here)
            assert exc is None
            exit_fn = pop()
            push(None)
            push(None)
            push(exit_fn(None, None, None))
        elif op == WITH_CLEANUP_FINISH:
            pop()
            pop()
        elif op == END_FINALLY:
            pop()
        else:
            raise NotImplementedError(f"py37vm: opcode {op} (arg {arg}) in {code.name} is not implemented")
    raise RuntimeError(f"py37vm: fell off the end of {code.name}")


def load_module(pyc_path: str, name: str = "reference_module") -> dict:
    """Execute a CPython-3.7 .pyc and return its global namespace."""
    import struct
    data = open(pyc_path, "rb").read()
    magic = struct.unpack_from("<H", data, 0)[0]
    if magic != 3394:
        raise ValueError(f"{pyc_path}: magic {magic} is not CPython 3.7 (3394)")
    top = Reader(data[16:]).obj()
    globs = {"__name__": name, "__builtins__": builtins, "__doc__": None}
    run_frame(top, globs, {})
    return globs
