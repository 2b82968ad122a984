"""Mnemonic counts per kernel from `cuobjdump -sass` of the in-tree extension -> profiles/sass_evidence.md.
Runs without a GPU (it only disassembles the sm_100a cubin embedded in the .so)."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sorted(glob.glob(os.path.join(ROOT, "sparktorch_b200", "_C*.so")))[0]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
COLS = [("UTCHMMA", "UTCHMMA (tcgen05.mma)"), ("UTMALDG", "UTMALDG (TMA load)"), ("LDTM", "LDTM (tcgen05.ld)"),
        ("UTMASTG", "UTMASTG (TMA store)"), ("\\.2CTA", ".2CTA (cta_group::2)"), ("IM2COL", "IM2COL (TMA im2col load)"),
        ("UTCBAR", "UTCBAR (tcgen05.commit)"), ("SYNCS", "SYNCS (mbarrier)"), ("MULTIMEM|LDGMC|STGMC|REDGMC", "multimem (NVLS)"),
        ("MEMBAR\\.\\S*SYS", "MEMBAR.SYS"), ("\\bRED\\b|\\bATOM|\\bATOMG|\\bREDG|\\bATOMS", "ATOM/RED"), ("ACQBULK|GRIDDEP|ACQBLK", "griddepcontrol"),
        ("SHFL", "SHFL"), ("\\bSTG", "STG")]
rows = []
cur, counts, n = None, None, 0
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, n, counts))
        cur, counts, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur:
        n += 1
        ins = m.group(1)
        for pat, name in COLS:
            if re.search(pat, ins):
                counts[name] += 1
if cur:
    rows.append((cur, n, counts))


def pretty(name: str) -> str:
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    out = re.sub(r"^void ", "", out).replace("(anonymous namespace)::", "")
    out = re.sub(r"\(.*", "", out)
    return out.replace("st::", "")


with open(os.path.join(ROOT, "profiles", "sass_evidence.md"), "w") as fh:
    fh.write("# SASS evidence (`cuobjdump -sass` of sparktorch_b200/_C*.so, sm_100a; regenerate with `python bench/sass_table.py`)\n\n")
    fh.write("Mnemonic counts per kernel.  `UTCHMMA` = `tcgen05.mma`, `UTMALDG` = TMA tensor load, `LDTM` = `tcgen05.ld`, "
             "`UTCBAR` = `tcgen05.commit`, `LDGMC` = `multimem.ld_reduce` (B200_PROFILING.md table).\n\n")
    fh.write("| kernel | instrs | " + " | ".join(n for _, n in COLS) + " |\n|---|---|" + "---|" * len(COLS) + "\n")
    for name, n, c in sorted(rows, key=lambda r: pretty(r[0])):
        fh.write(f"| `{pretty(name)[:70]}` | {n} | " + " | ".join(str(c.get(nm, 0)) for _, nm in COLS) + " |\n")
print(len(rows), "kernels")
