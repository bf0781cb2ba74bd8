"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md) in libwindsr.so.
    python scripts/sass_markers.py > profiles/r02_sass_markers.txt        (no GPU needed: cuobjdump reads the fatbin)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gan_sr_wind_field_b200", "libwindsr.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
marks = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.5D", "UTMALDG.3D", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UBLKCP", "REDG", "ATOMG"]
cur, cnt, arch = None, collections.OrderedDict(), set()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"ws::\(anonymous namespace\)::|ws::<unnamed>::", "", name)
        cur = re.sub(r"\(.*", "", name)
        cnt.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if cur:
        for k in marks:
            if re.search(r"\b" + re.escape(k) + r"(\.|\b)", line) and (k != "UTCHMMA" or True):
                cnt[cur][k] += 1
print(f"# SASS markers per kernel in gan_sr_wind_field_b200/libwindsr.so (cuobjdump -sass; arch {sorted(arch)})")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops")
print(f"{'kernel':64s} " + " ".join(f"{k:>12s}" for k in marks))
for k, c in cnt.items():
    if any(c.values()):
        print(f"{k[:64]:64s} " + " ".join(f"{c[m]:12d}" for m in marks))
tot = collections.Counter()
for c in cnt.values():
    tot.update(c)
print(f"{'TOTAL':64s} " + " ".join(f"{tot[m]:12d}" for m in marks))
