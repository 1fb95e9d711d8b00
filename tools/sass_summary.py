"""profiles/sass_summary.txt: per kernel of libb200fm.so, how many SASS instruction sites are tcgen05 / TMA (cuobjdump -sass, no GPU needed).
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "ml-4m_b200", "b200fm", "libb200fm.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
counts = collections.OrderedDict()
cur = None
KEYS = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "HMMA")
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in KEYS:
        if re.search(r"(?<![A-Z])" + k + r"(?![A-Z])", line):
            counts[cur][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS evidence per kernel of ml-4m_b200/b200fm/libb200fm.so (cuobjdump -sass, sm_100a); counts of instruction sites")
print("# UTCHMMA = tcgen05.mma (bf16 kind::f16), UTMALDG = TMA tensor load, UTMASTG = TMA tensor store (cp.async.bulk.tensor), LDTM = tcgen05.ld (TMEM -> registers),")
print("# HMMA = legacy mma.sync (must be 0)")
print(f"{'UTCHMMA':>8} {'UTMALDG':>8} {'UTMASTG':>8} {'LDTM':>6} {'HMMA':>5}  kernel (demangled)")
rows = []
for (mangled, c), name in zip(counts.items(), dem):
    name = re.sub(r"\(.*", "", name)
    if not any(c.values()) and "gemm" not in name and "attention" not in name:
        continue
    rows.append((name, c))
for name, c in sorted(rows):
    print(f"{c['UTCHMMA']:8d} {c['UTMALDG']:8d} {c['UTMASTG']:8d} {c['LDTM']:6d} {c['HMMA']:5d}  {name}")
print(f"# {len(counts)} kernels in the library; the ones without tensor / TMA instructions (element-wise, LayerNorm, gather, AdamW, all-reduce, gemv, ...) are omitted")
