#!/usr/bin/env python3
"""glsl2cpp.py — makes the reference's GLSL shader sources compilable as C++ (TEST INFRASTRUCTURE).

    glsl2cpp.py <reference>/src/shaders <out_dir> [--define NAME=VALUE ...]

Walks every .glsl/.comp/.frag file of the reference's shader directory and writes a token-level
transliteration of it to the same relative path under <out_dir> (oracle/_ref/glsl/, git-ignored: no
reference source enters the repository).  The arithmetic, control flow and operation order of every
function are left exactly as the authors wrote them; `#line` directives point back at the reference
file, so compiler diagnostics (and gdb) cite /root/reference/src/shaders/...:line.  What changes:

  1. unsuffixed floating literals get an `f` (GLSL `0.5` is a 32-bit float, C++ `0.5` is a double);
  2. `inout T x` / `out T x` parameters become `T &x`;
  3. `#version` / `#extension` lines are dropped;
  4. interface declarations (`layout(...) buffer|uniform|in|out ...;`) become plain C++ globals of the
     same names — `T name[];` members become `T *name;` — which oracle/ref_build/glsl_ref.cpp binds to
     the caller's arrays (the descriptor-set bindings of restirPass.h:120-237 etc.);
  5. GLSL evaluates function arguments left to right (spec §6.1.1), C++ leaves their order unspecified and
     g++ goes right to left: a statement with more than one `randFloat(...)` call gets the draws hoisted
     into temporaries, in source order, on the same line (restirOmni.glsl:111 and :125);
  6. optional `--define NAME=VALUE` rewrites the value of an existing `#define NAME ...` line (used for
     the north-star's 5-neighbour variant of unbiasedReuse.glsl:47 `#define NUM_NEIGHBORS 3` and for
     restirStructs.glsl:17 `#define RESERVOIR_SIZE 1`); `--enable NAME` uncomments a switch the authors ship
     commented out (restirStructs.glsl:16 `/*#define UNBIASED_MIS*/`).

GLSL types and built-ins (vec3, swizzles, dot, normalize, texelFetch, ...) come from glsl_shim.h.
"""
import os
import re
import sys

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
PARAM_REF = re.compile(r"\b(?:inout|out)\s+(\w+)\s+(\w+)")
LAYOUT = re.compile(r"layout\s*\([^)]*\)\s*(?P<rest>[^;{]*?)\s*(?:\{(?P<body>[^}]*)\}\s*(?P<inst>\w*)\s*)?;", re.S)
RAND_CALL = re.compile(r"\brandFloat\s*\(\s*(\w+)\s*\)")
MEMBER = re.compile(r"\s*([\w]+)\s+(\w+)\s*(\[\s*\w*\s*\])?\s*;")


def member_decl(m, static):
    typ, name, arr = m.group(1), m.group(2), m.group(3)
    pre = "static " if static else ""
    if arr is None:
        return f"{pre}{typ} {name};"
    if arr.strip("[] \t") == "":
        return f"{pre}{typ} *{name};"      # runtime-sized array of a storage block
    return f"{pre}{typ} {name}{arr};"


def translate_layout(m):
    rest, body, inst = m.group("rest").split(), m.group("body"), m.group("inst")
    src = m.group(0)
    keep_lines = "\n" * src.count("\n")
    if body is not None:
        members = [member_decl(x, static=not inst) for x in MEMBER.finditer(body)]
        if inst:
            return f"static struct {rest[1]}_block {{ {' '.join(members)} }} {inst};" + keep_lines
        return " ".join(members) + keep_lines
    if not rest or rest == ["in"]:                  # layout(local_size_x = ...) in;
        return keep_lines
    if "rayPayloadEXT" in rest or "accelerationStructureEXT" in rest:
        return keep_lines
    if rest[0] == "uniform":                        # uniform sampler2D name;
        return f"static {rest[1]} {rest[2]};" + keep_lines
    if rest[0] in ("in", "out"):                    # stage inputs / outputs: one value per invocation
        return f"static thread_local {rest[1]} {rest[2]};" + keep_lines
    raise SystemExit(f"glsl2cpp: unhandled interface declaration: {src!r}")


def code_part(line, fn):
    """Applies fn to the part of the line before a // comment."""
    i = line.find("//")
    return fn(line) if i < 0 else fn(line[:i]) + line[i:]


def sequence_draws(code, line_no):
    """Hoists the RNG draws of a statement that has several of them into temporaries, left to right."""
    calls = list(RAND_CALL.finditer(code))
    if len(calls) < 2:
        return code
    indent = re.match(r"\s*", code).group(0)
    names = [f"glsl_draw_{line_no}_{k}" for k in range(len(calls))]
    decl = "float " + ", ".join(f"{n} = {c.group(0)}" for n, c in zip(names, calls)) + "; "
    it = iter(names)
    return indent + decl + RAND_CALL.sub(lambda m: next(it), code).lstrip()


def translate(text, ref_path, defines, enables=()):
    for name in enables:
        text = re.sub(r"/\*\s*#\s*define\s+" + re.escape(name) + r"\s*\*/", "#define " + name, text)
    text = LAYOUT.sub(translate_layout, text)
    out = [f'#line 1 "{ref_path}"']
    for no, line in enumerate(text.split("\n"), 1):
        s = line.strip()
        if re.match(r"#\s*(version|extension)\b", s):
            out.append("")
            continue
        d = re.match(r"(\s*#\s*define\s+)(\w+)(\s+)(.*)$", line)
        if d and d.group(2) in defines:
            line = d.group(1) + d.group(2) + d.group(3) + defines[d.group(2)]
        if not re.match(r"#\s*include\b", s):
            line = code_part(line, lambda c: sequence_draws(PARAM_REF.sub(r"\1 &\2", FLOAT_LIT.sub(r"\1f", c)), no))
        out.append(line)
    return "\n".join(out) + "\n"


def main():
    args = sys.argv[1:]
    defines = {}
    while "--define" in args:
        i = args.index("--define")
        k, v = args[i + 1].split("=", 1)
        defines[k] = v
        del args[i:i + 2]
    enables = []
    while "--enable" in args:
        i = args.index("--enable")
        enables.append(args[i + 1])
        del args[i:i + 2]
    src_root, out_root = args
    n = 0
    for dirpath, _, files in os.walk(src_root):
        for f in files:
            if not f.endswith((".glsl", ".comp", ".frag")):
                continue
            p = os.path.join(dirpath, f)
            rel = os.path.relpath(p, src_root)
            q = os.path.join(out_root, rel)
            os.makedirs(os.path.dirname(q), exist_ok=True)
            with open(p) as fh:
                text = fh.read()
            with open(q, "w") as fh:
                fh.write(translate(text, os.path.abspath(p), defines, enables))
            n += 1
    print(f"glsl2cpp: {n} shader files -> {out_root}")


if __name__ == "__main__":
    main()
