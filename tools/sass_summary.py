"""profiles/r2_sass_summary.md: occurrences per kernel of the SASS mnemonics that identify the Blackwell paths
(`cuobjdump -sass` of the in-tree library; runs without a GPU).

    python tools/sass_summary.py > profiles/r2_sass_summary.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dotaclient_b200", "libdotaclient_b200.so")
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UCGABAR_ARV", "UCGABAR_WAIT", "SYNCS", "CCTL.E.PF2", "FFMA2", "HMMA", "MUFU.EX2",
        "STG.E.ENL2.256"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, fn = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            order.append(fn)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            for c in COLS:
                if op == c or op.startswith(c + ".") or (c in ("UCGABAR_ARV", "UCGABAR_WAIT", "UBLKCP", "UTCHMMA", "LDTM", "STTM") and op.startswith(c)):
                    counts[fn][c] += 1
    names = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
    print("# SASS instruction summary of `libdotaclient_b200.so` (round 2, committed kernels)\n")
    print("`cuobjdump -sass dotaclient_b200/libdotaclient_b200.so` (`tools/sass_summary.py`), occurrences per kernel of the mnemonics that identify the")
    print("Blackwell paths (`UTCHMMA` = tcgen05.mma kind::tf32, `LDTM`/`STTM` = tcgen05.ld/st, `UTCBAR` = tcgen05.commit, `UBLKCP` = cp.async.bulk (1-D TMA),")
    print("`UCGABAR_*` = barrier.cluster, `SYNCS` = mbarrier ops, `CCTL.E.PF2` = prefetch.global.L2, `FFMA2` = packed fp32x2 FMA; no `HMMA` (legacy mma.sync)")
    print("and no tensor-map TMA (`UTMALDG`) anywhere: operand staging is register-mediated because of the hi/lo split; compiled for sm_100a only).")
    print("Kernels without any of these mnemonics are omitted.\n")
    print("| kernel | " + " | ".join(COLS) + " |")
    print("|" + "---|" * (len(COLS) + 1))
    total = collections.Counter()
    for fn, name in zip(order, names):
        if not counts[fn]:
            continue
        short = re.sub(r"\(anonymous namespace\)::", "", name)
        short = re.sub(r"\(.*", "", short).replace("void ", "")
        print("| `%s` | " % short + " | ".join(str(counts[fn][c]) for c in COLS) + " |")
        total.update(counts[fn])
    print("| **total** | " + " | ".join(str(total[c]) for c in COLS) + " |")


if __name__ == "__main__":
    main()
