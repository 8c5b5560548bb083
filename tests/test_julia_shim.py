"""The Julia side of the boundary (julia/PIQLB200.jl) cannot run here -- there is no julia binary in this image or on
the GPU boxes -- so it is checked statically against include/piql_b200.h: every `ccall` names a declared entry point,
passes as many arguments as the prototype has, with Julia types that have the prototype's C layout, and every Julia
struct handed to the library mirrors the C struct field by field (and is isbits, so that Ref{T} / Vector{T} are the
C struct / C array)."""
import os
import re

from conftest import ROOT

HDR = os.path.join(ROOT, "include", "piql_b200.h")
JL = os.path.join(ROOT, "julia", "PIQLB200.jl")

# Julia struct <-> C struct
STRUCTS = {"GridParams": "piql_grid_params", "SolveOpts": "piql_solve_opts", "SolveStats": "piql_solve_stats",
           "McOpts": "piql_mc_opts", "McStats": "piql_mc_stats", "TableProblem": "piql_table_problem",
           "PeerHandles": "piql_peer_handles"}
SCALARS = {"double": {"Cdouble", "Float64"}, "float": {"Cfloat", "Float32"}, "int32_t": {"Int32", "Cint"},
           "int64_t": {"Int64"}, "uint64_t": {"UInt64"}, "uint8_t": {"UInt8"}, "uint32_t": {"UInt32"}}


def strip_c_comments(text):
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def split_top(s, sep=","):
    """split at separators that are not inside (), {} or []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def c_prototypes():
    hdr = strip_c_comments(open(HDR).read())
    protos = {}
    for ret, name, args in re.findall(r"\b(int32_t|const char\*)\s+(piql_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        params = [] if args.strip() == "void" else [re.sub(r"\s+", " ", a.strip()) for a in args.split(",")]
        # drop the parameter name, keep the type
        types = []
        for p in params:
            m = re.match(r"(.*?[\s\*])([A-Za-z_][A-Za-z0-9_]*)$", p)
            types.append(m.group(1).replace(" ", "") if m else p.replace(" ", ""))
        protos[name] = (ret, types)
    return protos


def c_structs():
    hdr = strip_c_comments(open(HDR).read())
    out = {}
    for body, name in re.findall(r"typedef struct\s*\{(.*?)\}\s*(piql_[a-z_]+)\s*;", hdr, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = re.sub(r"\s+", " ", decl.strip())
            if not decl:
                continue
            m = re.match(r"(.*?[\s\*])([A-Za-z_][A-Za-z0-9_\[\], ]*)$", decl)
            ctype = m.group(1).replace(" ", "")
            for nm in m.group(2).split(","):
                nm = nm.strip()
                arr = re.match(r"([A-Za-z_0-9]+)\[(\d+)\]", nm)
                fields.append((arr.group(1), ctype + "[%s]" % arr.group(2)) if arr else (nm, ctype))
        out[name] = fields
    return out


def julia_type_ok(jt, ct):
    """does the Julia ccall / field type jt have the layout of the C type ct?"""
    jt = jt.replace(" ", "")
    const_stripped = ct.replace("const", "")
    if const_stripped in SCALARS:
        return jt in SCALARS[const_stripped]
    arr = re.match(r"uint8_t\[(\d+)\]", const_stripped)
    if arr:
        return jt == "NTuple{%s,UInt8}" % arr.group(1)
    if const_stripped == "piql_ctx**":
        return jt == "Ref{Ptr{Cvoid}}"
    if const_stripped in ("piql_ctx*", "void*", "void**"):
        return jt in ("Ptr{Cvoid}", "Ptr{Ptr{Cvoid}}", "Ref{Ptr{Cvoid}}")
    if const_stripped == "char*":
        return jt == "Cstring"
    if const_stripped.endswith("*"):
        base = const_stripped[:-1]
        if base in SCALARS:
            return any(jt in ("Ptr{%s}" % j, "Ref{%s}" % j) for j in SCALARS[base])
        for jname, cname in STRUCTS.items():
            if base == cname:
                return jt in ("Ptr{%s}" % jname, "Ref{%s}" % jname)
    return False


def julia_ccalls(text):
    calls = []
    for m in re.finditer(r"ccall\(\(:(piql_[a-z0-9_]+),\s*LIB\),", text):
        # balanced scan from the opening parenthesis of ccall(
        start = text.index("(", m.start())
        depth, i = 0, start
        while True:
            depth += text[i] == "("
            depth -= text[i] == ")"
            if depth == 0:
                break
            i += 1
        parts = split_top(text[start + 1:i])
        name, ret, argt, args = m.group(1), parts[1], parts[2], parts[3:]
        assert argt.startswith("(") and argt.endswith(")"), (name, argt)
        types = split_top(argt[1:-1])
        calls.append((name, ret, types, args))
    return calls


def julia_structs(text):
    text = re.sub(r"#.*", "", text)
    out = {}
    for mut, name, body in re.findall(r"^(mutable\s+)?struct\s+(\w+)\s*;?(.*?)\bend\b", text, flags=re.S | re.M):
        fields = re.findall(r"(\w+)::([\w{},]+)", body)
        out[name] = (bool(mut), fields)
    return out


def test_every_ccall_matches_its_prototype():
    protos = c_prototypes()
    assert len(protos) > 40, "header prototypes did not parse"
    calls = julia_ccalls(open(JL).read())
    assert len(calls) >= 20, "ccalls did not parse"
    for name, ret, types, args in calls:
        assert name in protos, "%s is not declared in the header" % name
        cret, ctypes_ = protos[name]
        assert ret == ("Cstring" if cret == "const char*" else "Int32"), (name, ret)
        assert len(types) == len(ctypes_), "%s: %d Julia argument types, prototype has %d" % (name, len(types), len(ctypes_))
        assert len(args) == len(types), "%s: %d arguments for %d types" % (name, len(args), len(types))
        for k, (jt, ct) in enumerate(zip(types, ctypes_)):
            assert julia_type_ok(jt, ct), "%s argument %d: Julia %s vs C %s" % (name, k, jt, ct)


def test_the_shim_covers_the_public_solver_surface():
    called = {c[0] for c in julia_ccalls(open(JL).read())}
    need = {"piql_create", "piql_destroy", "piql_last_error", "piql_grid_load", "piql_solve_nu", "piql_solve_logz",
            "piql_update_v", "piql_solve_gap", "piql_solve_z", "piql_z_update_v", "piql_mc_logz", "piql_mc_update_v",
            "piql_path_estimates", "piql_learning_traj", "piql_solve_gap_batch", "piql_table_load", "piql_table_solve",
            "piql_peer_export", "piql_peer_import", "piql_mc_allreduce_peer"}
    assert need <= called, need - called


def test_julia_structs_mirror_the_c_structs():
    cs = c_structs()
    js = julia_structs(open(JL).read())
    for jname, cname in STRUCTS.items():
        assert jname in js and cname in cs, (jname, cname)
        mutable, jfields = js[jname]
        # a mutable struct is a boxed object: a Vector of them is an array of pointers, not the C array the batched
        # entry points write (piql_learning_traj / piql_solve_gap_batch stats_out)
        assert not mutable, "%s must be an immutable (isbits) struct" % jname
        cfields = cs[cname]
        assert [f for f, _ in jfields] == [f for f, _ in cfields], (jname, jfields, cfields)
        for (fn, jt), (_, ct) in zip(jfields, cfields):
            if ct.replace("const", "").endswith("*"):
                base = ct.replace("const", "")[:-1]
                assert jt in {"Ptr{%s}" % j for j in SCALARS[base]}, (jname, fn, jt, ct)
            else:
                assert julia_type_ok(jt, ct), (jname, fn, jt, ct)


def test_batched_stats_buffers_are_isbits_vectors():
    text = open(JL).read()
    # the two batched calls hand a Vector{SolveStats} to a `piql_solve_stats*` that the library indexes as an array
    assert len(re.findall(r"Vector\{SolveStats\}\(undef", text)) == 2
    assert "mutable struct SolveStats" not in text and "mutable struct McStats" not in text


# The reference's solver signatures (src/value_iteration.jl:42,54,77,107,119,144 -- `function generate_ν(ctrl)`,
# `generate_logz(ctrl, V; β = 1.0)`, `logz_updateV(ctrl, V; β = 1.0)`, `generate_z(ctrl, V; β = 1.0)`,
# `z_updateV(ctrl, V; β = 1.0)`, `generate_optimality_gap(ctrl, V; ν = generate_ν(ctrl))`): the shim must extend exactly
# these functions of PIQL with methods on ControlProblem, so scripts/experiments.jl:10-39 runs unchanged.
REFERENCE_SIGNATURES = {
    "generate_ν": ([], []),
    "generate_logz": (["V"], ["β = 1.0"]),
    "logz_updateV": (["V"], ["β = 1.0"]),
    "generate_z": (["V"], ["β = 1.0"]),
    "z_updateV": (["V"], ["β = 1.0"]),
    "generate_optimality_gap": (["V"], ["ν = generate_ν(ctrl)"]),
}


def test_reference_signatures_are_extended_on_controlproblem():
    text = open(JL).read()
    imp = re.search(r"^import PIQL: (generate_ν.*)$", text, flags=re.M)
    assert imp, "the solver functions must be imported from PIQL to be extended (not shadowed)"
    imported = {n.strip() for n in imp.group(1).split(",")}
    assert set(REFERENCE_SIGNATURES) <= imported, set(REFERENCE_SIGNATURES) - imported
    for name, (pos, kws) in REFERENCE_SIGNATURES.items():
        m = re.search(r"^function %s\(ctrl::ControlProblem(.*?)\)\s*$" % re.escape(name), text, flags=re.M)
        assert m, "no method %s(ctrl::ControlProblem, ...)" % name
        rest = m.group(1)
        positional, _, keywords = rest.partition(";")
        got_pos = [a.strip() for a in positional.split(",") if a.strip()]
        assert got_pos == pos, (name, got_pos)
        got_kw = [k.strip() for k in split_top(keywords) if k.strip()]
        for k in kws:      # the reference's keywords with the reference's defaults; extra solver options ride in kw...
            assert k in got_kw, (name, k, got_kw)
        assert got_kw[-1] == "kw..."
    # the registry recognises gridworld_ctrl problems from the closures' captured variables and reloads on change
    assert "function gridworld_of(ctrl::ControlProblem)" in text and "function context_for(ctrl::ControlProblem)" in text
    assert re.search(r"key = \(:grid, hash\(gw\.walls\)", text), "the loaded-problem key must hash the wall contents"
    # scripts/experiments.jl:20 spells run_experiment(gw; ...): no context argument
    assert re.search(r"^run_experiment\(gw; kw\.\.\.\) =", text, flags=re.M)
