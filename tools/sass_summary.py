"""SASS evidence per kernel of libsmelter_b200.so: tcgen05 / TMA / TMEM mnemonics counted with cuobjdump (no GPU needed).
Usage: python tools/sass_summary.py > profiles/rN_sass_summary.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "smelter_b200", "libsmelter_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
chunks = re.split(r"\n\s*Function : \S+", sass)[1:]
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "IM2COL", "UTMASTG", "UTMAPF", "SYNCS", "UCGABAR", "HMMA", "STG.E.128", "LDG.E.128"]
print(f"# cuobjdump -sass smelter_b200/libsmelter_b200.so (sm_100a): instruction mnemonics per kernel ({len(chunks)} kernels)")
print("# " + "  ".join(KEYS) + "  | kernel")
tot = {k: 0 for k in KEYS}
for name, body in zip(names, chunks):
    counts = []
    for k in KEYS:
        if k == "UTCHMMA":
            n = len(re.findall(r"\bUTCHMMA\b(?!\.2CTA)", body))
        elif k == "HMMA":
            n = len(re.findall(r"\sHMMA\.", body))
        else:
            n = body.count(k)
        counts.append(n)
        tot[k] += n
    short = re.sub(r"smelter::k::\(anonymous namespace\)::", "", name)
    short = re.sub(r"\(CUtensorMap_st.*", "(...)", short)[:110]
    print("  ".join(f"{c:>{len(k)}d}" for c, k in zip(counts, KEYS)) + "  | " + short)
print("# totals: " + ", ".join(f"{k} {v}" for k, v in tot.items()))
