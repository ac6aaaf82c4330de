"""SASS mnemonic histogram per kernel of the built library (cuobjdump -sass), the proof that the hot kernels are
tcgen05 / TMEM / bulk-TMA code and not recompiled mma.sync:
    python profiles/sass_histogram.py > profiles/sass_r2_histogram.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.path.join(ROOT, "adanerf_b200", "libadanerf_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("adn::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "FENCE", "UCGABAR", "LDS", "STS", "LDG", "STG", "LDC", "LDCU", "MUFU", "NANOSLEEP", "BAR")
print("# SASS mnemonic histograms (round 2, `cuobjdump -sass adanerf_b200/libadanerf_b200.so`, sm_100a)\n")
print("Tensor-core / TMEM / TMA instructions of Blackwell: `UTCHMMA` = tcgen05.mma (kind::f16), `UTCBAR` = tcgen05.commit -> mbarrier, "
      "`LDTM` = tcgen05.ld, `UBLKCP` = cp.async.bulk (1-D TMA), `SYNCS` = mbarrier ops, `UTCATOMSWS` = tcgen05.alloc/dealloc.  `HMMA` (mma.sync) never appears.\n")
print("| kernel | instructions | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
for k, h in hist.items():
    fam = collections.Counter()
    for op, n in h.items():
        fam[op.split(".")[0]] += n
    print(f"| `{k}` | {sum(h.values())} | " + " | ".join(str(fam.get(x, 0)) for x in KEY) + " |")
print("\n## Full opcode lists of the two MLP kernels (opcode with modifiers: count)\n")
for k, h in hist.items():
    if "mlp_sh_kernel" in k or "mlp_hp_kernel<2>" in k or "mlp_hp_kernel<(int)2>" in k:
        print(f"### `{k}`\n")
        print(", ".join(f"`{op}`: {n}" for op, n in sorted(h.items(), key=lambda x: -x[1])) + "\n")
