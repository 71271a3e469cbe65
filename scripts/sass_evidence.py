"""Regenerate profiles/r2_sass_evidence.txt: per kernel of the shipped library, counts of the Blackwell-specific SASS mnemonics."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "catgrasp_b200", "lib", "libcatgrasp_b200.so")
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "SYNCS", "FMNMX3", "CREDUX", "REDUX", "UCGABAR", "USETMAXREG",
        "FFMA2", "FADD2", "F2FP", "LDGSTS", "SHFL", "LDS", "STS", "ATOMS", "LD.E", "ST.E", "ATOM.E", "MEMBAR"]
SHOW = ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UCGABAR", "F2FP.SATFINITE")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    out = ["# SASS evidence, catgrasp_b200/lib/libcatgrasp_b200.so (cuobjdump -sass; release build, sm_100a; scripts/sass_evidence.py)", "",
           "Per kernel: counts of the Blackwell-specific mnemonics (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,",
           "cp.async.bulk -> UBLKCP, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, 3-input max -> FMNMX3, packed fp32 -> FFMA2/FADD2, redux ->",
           "CREDUX/REDUX) and of shared / generic memory instructions; first occurrences of the tensor-memory / bulk-copy instructions quoted.", ""]
    cur, body = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            body[cur].append(line.rstrip())
    for fn, lines in body.items():
        cnt = collections.Counter()
        for l in lines:
            ins = l.split("*/", 1)[1].strip() if "*/" in l else ""
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            op = ins.split(" ")[0] if ins else ""
            for w in WANT:
                if op == w or op.startswith(w + ".") or (w in ("LD.E", "ST.E", "ATOM.E") and op.startswith(w)):
                    cnt[w] += 1
                    break
        if not any(cnt[w] for w in ("UTCHMMA", "LDTM", "UBLKCP", "UCGABAR", "CREDUX", "FMNMX3", "SYNCS")):
            continue
        out.append(f"## {fn}")
        out.append("  " + ", ".join(f"{w} x{cnt[w]}" for w in WANT if cnt[w]))
        shown = 0
        for l in lines:
            if any(s in l for s in SHOW) and shown < 6:
                out.append("    " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
                shown += 1
        out.append("")
    dst = os.path.join(ROOT, "profiles", "r2_sass_evidence.txt")
    open(dst, "w").write("\n".join(out))
    print(dst, len(out), "lines")


if __name__ == "__main__":
    main()
