"""Which kernels of the shipped libecgmm.so use the Blackwell tensor-core / TMA instructions, counted from the SASS.

    python tools/sass_census.py > profiles/r02_sass_tcgen05.txt

Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma kind::f16 (".2CTA" = cta_group::2), UTCBAR = tcgen05.commit,
UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / tcgen05.st, SYNCS = mbarrier operations."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "ecg-multimodal-model_b200", "libecgmm.so")
PAT = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCBAR", r"\bUTCBAR"),
       ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
       ("SYNCS", r"\bSYNCS")]


def main():
    text = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        for name, pat in PAT:
            if re.search(pat, line):
                counts[cur][name] += 1
    names = list(counts)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)}: instruction counts per kernel (static, not executed counts); "
          f"{len(names)} kernels, {sum(1 for n in names if counts[n]['UTCHMMA'] or counts[n]['UTCHMMA.2CTA'])} with tcgen05.mma")
    print(f"{'kernel':64s} " + " ".join(f"{n:>12s}" for n, _ in PAT))
    rows = []
    for n, d in zip(names, dem):
        c = counts[n]
        if not any(c[k] for k, _ in PAT if k != "SYNCS"):
            continue
        short = d.split("(")[0].replace("ecgmm::", "").replace("void ", "")
        rows.append((short, c))
    for short, c in sorted(rows):
        print(f"{short[:64]:64s} " + " ".join(f"{c[k]:12d}" for k, _ in PAT))


if __name__ == "__main__":
    main()
