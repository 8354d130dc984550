#!/usr/bin/env python
"""Where a kernel's warp instructions go, from the source page of an `ncu --set full --import-source on` report.
Usage: python tools/ncu_source_regions.py <report.ncu-rep> <launch index>
Prints the opcode mix, the share of instructions executed behind the kernel's EXIT (subroutines: the divider's slow path), the CALL
sites that are actually taken, and the SASS regions of (nearly) equal execution count with their lanes per instruction and share of
the stall samples."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, k = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", k, "--launch-count", "1"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    print(rows[0][1])
    idx = {h: i for i, h in enumerate(rows[1])}
    lines = []
    for r in rows[2:]:
        try:
            n = int(r[idx["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        src = r[idx["Source"]]
        op = src.split()[1] if src.startswith("@") else src.split()[0]
        lines.append((r[idx["Address"]][-5:], op, n, int(r[idx["Thread Instructions Executed"]] or 0), int(r[idx["# Samples"]] or 0)))
    tot, ts = sum(l[2] for l in lines), max(sum(l[4] for l in lines), 1)
    mix = collections.Counter()
    for _, op, n, _, _ in lines:
        mix[op.split(".")[0]] += n
    print("warp instructions %.1f M; " % (tot / 1e6) + " ".join("%s %.1f%%" % (o, 100.0 * n / tot) for o, n in mix.most_common(16)))
    behind, seen = 0, False
    for _, op, n, _, _ in lines:
        if seen:
            behind += n
        seen = seen or op == "EXIT"
    print("behind EXIT: %.1f %%" % (100.0 * behind / tot))
    for a, op, n, t, _ in lines:
        if op.startswith("CALL") and n > tot * 0.0003:
            print("  CALL at %s taken %.3f M times, %.1f lanes" % (a, n / 1e6, t / max(n, 1)))
    cur, runs = None, []
    for a, op, n, t, s in lines:
        if cur and n > 0 and cur["n0"] > 0 and 0.7 < n / cur["n0"] < 1.4:
            cur["n"] += n; cur["t"] += t; cur["s"] += s; cur["k"] += 1; cur["end"] = a; cur["ops"].append(op)
        else:
            if cur:
                runs.append(cur)
            cur = {"start": a, "end": a, "n0": n, "n": n, "t": t, "s": s, "k": 1, "ops": [op]}
    runs.append(cur)
    for r in runs:
        if r["n"] > tot * 0.005:
            print("%s-%s %4d instr x %7.3f M = %5.1f %%  lanes %4.1f  samples %5.1f %%  %s" % (
                r["start"], r["end"], r["k"], r["n"] / r["k"] / 1e6, 100.0 * r["n"] / tot, r["t"] / max(r["n"], 1), 100.0 * r["s"] / ts, " ".join(r["ops"][:6])))


if __name__ == "__main__":
    main()
