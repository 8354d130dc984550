#!/usr/bin/env python
"""Per-kernel SASS summary of librestir_b200.so (cuobjdump, no GPU needed): registers, instruction count and the opcodes the
design leans on — packed FP32 (FADD2 / FMUL2 / FFMA2), 256-bit loads (LDG.E...256), MUFU, 64-bit atomics, shuffles, votes —
plus an excerpt of the trace kernel's node visit.  Usage: python profiles/sass_histogram.py [lib.so] > profiles/r2_sass_summary.md"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "restir-vulkan_b200/librestir_b200.so"
KEYS = ["FADD2", "FMUL2", "FFMA2", "PRMT", "LDG.E.ENL2.256", "LDG.E.128", "LDG.E.64", "LDG.E ", "LDS", "STS", "LDL", "STL", "MUFU", "FMNMX3", "FMNMX ", "SHFL", "VOTE",
        "MATCH", "ATOM", "RED", "BAR", "HMMA", "UTMA", "TCGEN05"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and cur:
            regs[cur] = tuple(int(v) for v in m.groups())
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    arch = set()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            kernels[cur].append(m.group(1).strip())
    print("# SASS summary of librestir_b200.so (cuobjdump -sass / -res-usage, CUDA 12.9, built by restir-vulkan_b200/build.py)\n")
    print(f"Architectures in the fat binary: {sorted(arch)} — sm_100a cubins only, no PTX fallback for other targets.\n")
    print("No tensor-core (HMMA / tcgen05) and no TMA (UTMA) instruction anywhere: nothing on this path is a dense contraction and every "
          "access is a per-lane gather or a per-pixel record (DESIGN.md §4).  Packed FP32 (`FADD2` / `FMUL2`, two IEEE binary32 operations per "
          "issue slot, each rounded like the scalar one) carries the slab tests of the binary walk, packed `FFMA2` + `PRMT` the conservative box "
          "tests of the 4-wide walk (`trace_wide_kernel`, csrc/restir_wide.cuh); 256-bit loads fetch the 64-byte nodes of both.\n")
    cols = ["kernel", "regs", "smem B", "local B", "instr"] + [k.strip() for k in KEYS]
    print("| " + " | ".join(cols) + " |")
    print("|" + "---|" * len(cols))
    for name, ins in kernels.items():
        pretty = demangle(name)
        pretty = re.sub(r"\(.*", "", pretty).replace("restir::", "").replace("void ", "")
        r = regs.get(name, (0, 0, 0))
        counts = [sum(1 for i in ins if re.sub(r"^@!?U?P\d+\s+", "", i).startswith(k)) for k in KEYS]
        print(f"| `{pretty}` | {r[0]} | {r[1]} | {r[2]} | {len(ins)} | " + " | ".join(str(c) if c else "" for c in counts) + " |")
    # the node visit of the wide walk (pixel mode) and of the binary walk
    for name, ins in kernels.items():
        if "trace_wide_kernel" in name and "Li0E" in name:
            loads = [i for i, s in enumerate(ins) if "LDG.E.ENL2.256" in s]
            first = loads[1]  # [0] is the triangle record of the cache pretest
            print("\n## `trace_wide_kernel<pixel>`: one node visit (four quantised boxes of a 64-byte node)\n\n```")
            print("\n".join(ins[first - 10:first + 66]))
            print("```")
            break
    for name, ins in kernels.items():
        if "trace_kernel" in name and "Li0ELi1" in name:
            first = next(i for i, s in enumerate(ins) if "LDG.E.ENL2.256" in s)
            print("\n## `trace_kernel<pixel, image>`: one node visit of the binary walk (both boxes of a 64-byte node)\n\n```")
            print("\n".join(ins[first - 2:first + 46]))
            print("```")
            break


if __name__ == "__main__":
    main()
