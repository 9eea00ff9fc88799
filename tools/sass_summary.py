"""profiles/sass_summary.txt: which Blackwell instructions the SHIPPED library contains.

    python tools/sass_summary.py            (no GPU needed: cuobjdump on the built .so)

Counts SASS mnemonics per kernel (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor
load, UBLKCP = cp.async.bulk, UTMAPF = TMA L2 prefetch, UTCBAR = tcgen05.commit, STAS = st.async into a peer
CTA's shared memory, UCGABAR = cluster barrier, SYNCS = mbarrier operations) and the PTX-level
`tcgen05.` / `cp.async.bulk` strings in the sources."""
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "knn_svc_b200" / "libknnsvc_b200.so"
PAT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "UTCBAR", "HMMA", "SYNCS",
       "STAS", "UCGABAR", "DFMA", "FFMA2", "REDUX", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = defaultdict(Counter)
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn)
            continue
        if fn is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for p in PAT:
                if op.startswith(p):
                    per[fn][p] += 1
                    if p in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "UBLKCP", "UTMAPF"):
                        per[fn]["variant:" + op] += 1
    out = [f"SASS mnemonic counts in {LIB.name} (cuobjdump -sass, sm_100a), per kernel; only kernels with a hit are listed", ""]
    for fn in sorted(per):
        c = per[fn]
        main_ = ", ".join(f"{p}={c[p]}" for p in PAT if c[p])
        var = ", ".join(sorted(k[8:] for k in c if k.startswith("variant:")))
        out.append(f"{fn}\n    {main_}\n    forms: {var}" if var else f"{fn}\n    {main_}")
    out.append("")
    out.append("PTX-level strings in the sources (grep -c):")
    for src in sorted((ROOT / "knn_svc_b200" / "csrc").glob("*.cu")):
        t = src.read_text()
        n1, n2, n3 = len(re.findall(r"tcgen05\.", t)), len(re.findall(r"cp\.async\.bulk\.tensor", t)), len(
            re.findall(r"cp\.async\.bulk\.(?!tensor|prefetch)", t))
        if n1 or n2 or n3:
            out.append(f"  {src.name}: tcgen05.* x{n1}, cp.async.bulk.tensor x{n2}, cp.async.bulk (non-tensor) x{n3}")
    text = "\n".join(out) + "\n"
    (ROOT / "profiles" / "sass_summary.txt").write_text(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
