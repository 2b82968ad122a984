#!/usr/bin/env python
"""Per-phase timeline of one K-STEP step from the kernel's globaltimer stamps (SPARKTORCH_MEGA_PROF=<file> during a fit).

stamps per task: 0 wait start, 1 early deps met (B operand issue), 2 all deps met (A operand issue), 3 first operand stage
landed (MMA thread), 4 all MMAs issued, 5 accumulator complete (epilogue), 6 epilogue stores issued, 7 signalled."""
import collections
import json
import sys


def main(path, md=None):
    p = json.load(open(path))
    ts = [t for t in p["tasks"] if any(t["t"])]
    t0 = min(min(x for x in t["t"] if x) for t in ts)
    g = collections.OrderedDict()
    for t in p["tasks"]:
        g.setdefault(t["group"], []).append([(x - t0) / 1000 if x else None for x in t["t"]])
    names = ["wait", "earlyB", "deps", "1st-stage", "mma-issued", "acc-ready", "stores", "signalled"]
    lines = ["| group | tasks | " + " | ".join(names) + " |", "|---|---|" + "---|" * len(names)]
    for k, v in g.items():
        cells = []
        for i in range(8):
            xs = [s[i] for s in v if s[i] is not None]
            cells.append("%.1f-%.1f" % (min(xs), max(xs)) if xs else "")
        lines.append(f"| {k} | {len(v)} | " + " | ".join(cells) + " |")
    end = max(x for t in ts for x in t["t"] if x)
    out = [f"K-STEP timeline (us from the first stamp; world={p.get('world')}, {p['info']}); step span {(end - t0) / 1000:.1f} us", ""] + lines
    # per-k-block view of the first task of a few groups: B issued / A issued / stage landed, relative to the task's "deps"
    for grp in ("F1", "F2", "B2", "G1"):
        for t in p["tasks"]:
            if t["group"] == grp and t.get("kb"):
                base = t["t"][0]
                rows = ["%d:%s" % (i, "/".join("%.2f" % ((x - base) / 1000) if x else "-" for x in kb[:3])) for i, kb in enumerate(t["kb"])]
                out.append("")
                out.append(f"{grp} cta {t['cta']} k-blocks (B issued/A issued/landed, us after the task's wait start; deps at "
                           f"{(t['t'][2] - base) / 1000:.2f}): " + "  ".join(rows))
                break
    print("\n".join(out))
    if md:
        open(md, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
