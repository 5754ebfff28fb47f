"""The Julia shim cannot be executed in this image (no Julia), so its `ccall`s are checked statically against
include/cude_b200.h: every called symbol is declared, the number of argument types equals the prototype's parameter count,
pointer / scalar kinds agree position by position, and the two mirrored structs have the header's fields in order."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def _header_prototypes():
    src = _strip_comments(open(os.path.join(ROOT, "include", "cude_b200.h")).read())
    protos = {}
    for m in re.finditer(r"\b(?:int|void|const char\*|double)\s+(cude_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        params = [p.strip() for p in m.group(2).split(",")]
        params = [] if params in ([""], ["void"]) else params
        protos[m.group(1)] = params
    return src, protos


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _ccalls(path):
    src = re.sub(r"#[^\n]*", " ", open(path).read())
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*libcude\)\s*,", src):
        depth, i = 1, src.index("(", m.start()) + 1
        start = i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        args = _split_top(src[start:i - 1])
        # args[0] = (:name, libcude), args[1] = return type, args[2] = (types...), rest = values
        types = _split_top(args[2].strip()[1:-1]) if args[2].strip() != "()" else []
        types = [t for t in types if t]
        calls.append((m.group(1), args[1].strip(), types, args[3:]))
    return calls


def _is_pointer_c(p):
    return "*" in p


def _is_pointer_jl(t):
    return t.startswith(("Ptr{", "Ref{")) or t in ("Cstring",)


def test_every_ccall_matches_the_header():
    _, protos = _header_prototypes()
    assert len(protos) > 40
    n = 0
    for f in ("CUDEB200.jl", "cude_overrides.jl"):
        for name, ret, types, values in _ccalls(os.path.join(ROOT, "julia", f)):
            assert name in protos, f"{f}: {name} is not declared in include/cude_b200.h"
            params = protos[name]
            assert len(types) == len(params), f"{f}: {name} passes {len(types)} argument types, the header has {len(params)}: {params}"
            assert len(values) == len(types), f"{f}: {name} has {len(values)} values for {len(types)} types"
            for t, p in zip(types, params):
                assert _is_pointer_jl(t) == _is_pointer_c(p), f"{f}: {name}: Julia type {t} against C parameter '{p}'"
                if not _is_pointer_c(p):
                    base = p.split()[0] if not p.startswith("long long") else "long long"
                    want = {"int": "Cint", "double": "Cdouble", "long long": "Clonglong"}[base]
                    assert t == want, f"{f}: {name}: Julia type {t} against C parameter '{p}'"
            n += 1
    assert n >= 28


def test_mirrored_structs_have_the_headers_fields_in_order():
    src, _ = _header_prototypes()
    jl = open(os.path.join(ROOT, "julia", "CUDEB200.jl")).read()
    jl = re.sub(r"#[^\n]*", "", jl)
    for cname, jname in (("cude_opts", "CudeOpts"), ("cude_net", "CudeNet"), ("cude_train_opts", "CudeTrainOpts")):
        body = re.search(r"typedef struct(?:\s+\w+)?\s*\{([^}]*)\}\s*" + cname + r"\s*;", src).group(1)
        cfields = []
        for m in re.finditer(r"\b(int|double|long long)\s+([\w\s,]+);", body):          # `double a, b, c;` declares three fields
            cfields += [(m.group(1), n.strip()) for n in m.group(2).split(",")]
        jbody = re.search(r"struct " + jname + r"\n(.*?)\nend", jl, flags=re.S).group(1)
        jbody = jbody.split("\n    " + jname + "()")[0]                                   # drop an inner constructor
        jfields = [(m.group(2), m.group(1)) for m in re.finditer(r"^\s*(\w+)::(\w+)", jbody, flags=re.M)]
        assert [n for _, n in cfields] == [n for _, n in jfields], (cfields, jfields)
        kinds = {"int": "Cint", "double": "Cdouble", "long long": "Clonglong"}
        assert [kinds[t] for t, _ in cfields] == [t for t, _ in jfields], (cfields, jfields)
