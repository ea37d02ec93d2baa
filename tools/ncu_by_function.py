"""Development aid: aggregate an ncu report's SASS-level samples per device function.

    python tools/ncu_by_function.py gpurun_out/prof.ncu-rep [kernel-substring] [lib.so]

Uses `ncu --page source --csv` for per-instruction counters and the cubin's symbol table (cuobjdump
-xelf + readelf) to attribute each SASS address to the (non-inlined) device function containing it.
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep = sys.argv[1]
    kern_sub = sys.argv[2] if len(sys.argv) > 2 else "GaussPeaks"
    so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                            "enlsip.jl_b200", "lib", "libenlsip_b200.so")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[hdr_i], [r for r in rows[hdr_i + 1:] if r and r[0].startswith("0x")]
    col = {h: i for i, h in enumerate(hdr)}
    addrs = [int(r[col["Address"]], 16) for r in data]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    secs = subprocess.run(["readelf", "-SW", cubin], capture_output=True, text=True).stdout
    secidx = None
    for line in secs.split("\n"):
        m = re.match(r"\s*\[\s*(\d+)\]\s+(\S+)", line)
        if m and m.group(2).startswith(".text.") and kern_sub in m.group(2):
            secidx = m.group(1)
    syms = []
    for line in subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout.split("\n"):
        m = re.match(r"\s*\d+:\s+([0-9a-f]+)\s+(\S+)\s+FUNC\s+\S+\s+\S+(?:\s+\[<other>: \w+\])?\s+(\d+)\s+(\S+)", line)
        if m and m.group(3) == secidx:
            syms.append((int(m.group(1), 16), int(m.group(2), 0), m.group(4)))
    syms.sort()
    kern = [s for s in syms if "$" not in s[2]][0]
    syms = [s for s in syms if "$" in s[2]] + [kern]     # the kernel symbol spans the whole section: match it last
    base = addrs[0] - kern[0]

    def fn(a):
        off = a - base
        for v, sz, name in syms:
            if v <= off < v + sz:
                return name
        return "?"

    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    for r, a in zip(data, addrs):
        g = agg[fn(a)]
        g["inst"] += int(r[col["Instructions Executed"]] or 0)
        g["samples"] += int(r[col["# Samples"]] or 0)
        for h in stall_cols:
            g[h] += int(r[col[h]] or 0)
    tot_i = sum(v["inst"] for v in agg.values())
    tot_s = sum(v["samples"] for v in agg.values())

    def dem(n):
        n2 = n.split("$")[-1]
        d = subprocess.run(["c++filt", n2], capture_output=True, text=True).stdout.strip()
        d = re.sub(r"enl::Solver<.*?\d+>::", "S::", d)
        d = re.sub(r"enl::Dist<.*?>::", "Dist::", d)
        d = re.sub(r"\(.*", "", d)
        return d[:60]

    tot = collections.Counter()
    for v in agg.values():
        tot.update(v)
    print("total warp-instructions %d, samples %d, code bytes %d" % (tot_i, tot_s, sum(s[1] for s in syms)))
    print("stall mix:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * tot[h] / tot_s) for h in sorted(stall_cols, key=lambda h: -tot[h])[:7]))
    print("%7s %8s %7s  %-9s %s" % ("inst%", "samples%", "bytes", "top stall", "function"))
    size = {name: sz for _, sz, name in syms}
    for f, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:40]:
        top = max(stall_cols, key=lambda h: v[h])
        print("%6.1f%% %7.1f%% %7d  %-9s %s" % (100.0 * v["inst"] / tot_i, 100.0 * v["samples"] / tot_s, size.get(f, 0),
                                               "%s %.0f%%" % (top[6:14], 100.0 * v[top] / max(1, v["samples"])), dem(f)))


if __name__ == "__main__":
    main()
