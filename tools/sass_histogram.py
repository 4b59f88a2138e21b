"""Instruction histogram of the built library's SASS: per kernel, the counts of the mnemonics that prove the Blackwell data path
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, LDTM = tcgen05.ld, UTCBAR =
tcgen05.commit, SYNCS = mbarrier ops).  usage: python tools/sass_histogram.py [out.md]   (needs the .o files of a build)"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "adversarial-attacks-on-gan-based-image-fusion_b200", "csrc")
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "LDG", "STG", "LDS", "STS", "FFMA", "FFMA2", "FADD2", "FMUL2", "REDG", "SHFL"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    names = []
    for o in out:
        o = re.sub(r"\(anonymous namespace\)::|void |<unnamed>::|\(bool\)|\(int\)", "", o)
        names.append(o[:o.rfind("(")] if o.endswith(")") else o)      # drop the parameter list, keep template arguments
    return names


def main():
    rows = []
    for obj in sorted(f for f in os.listdir(CSRC) if f.endswith(".o")):
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
        cur, cnt, total = None, None, 0
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if cur:
                    rows.append((obj, cur, cnt, total))
                cur, cnt, total = m.group(1), collections.Counter(), 0
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and cur:
                total += 1
                op = m.group(1)
                for k in KEYS:
                    if op == k or op.startswith(k + "."):
                        cnt[k] += 1
        if cur:
            rows.append((obj, cur, cnt, total))
    names = demangle([r[1] for r in rows])
    lines = ["# SASS instruction histogram of libsfattack.so (sm_100a), per kernel", "",
             "`python tools/sass_histogram.py` over the object files of `python __graft_entry__.py build` (cuobjdump -sass).",
             "UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,",
             "SYNCS = mbarrier operations.  Kernels without any of the first six columns are plain CUDA-core kernels.", "",
             "| object | kernel | instr | " + " | ".join(KEYS) + " |", "|---|---|---:|" + "---:|" * len(KEYS)]
    tot = collections.Counter()
    for (obj, _, cnt, total), nm in zip(rows, names):
        lines.append(f"| {obj} | `{nm[:90]}` | {total} | " + " | ".join(str(cnt[k]) if cnt[k] else "" for k in KEYS) + " |")
        tot.update(cnt)
    lines.append("| **all** | | | " + " | ".join(f"**{tot[k]}**" for k in KEYS) + " |")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text[-600:])


if __name__ == "__main__":
    main()
