"""Attribute ncu warp-stall samples to CUDA source lines: joins `ncu --page source --csv` (SASS rows, in address order)
with `nvdisasm --print-line-info` of the object that was profiled.
  python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <object .o> <source .cu> [top N]
"""
import csv, io, re, subprocess, sys, tempfile, os, collections

def sass_rows(rep, kern):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
    return rows[hi], rows[hi + 1:]

def line_table(obj, kern):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    lines, cur, on = [], None, False
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            on = re.search(kern, ln) is not None
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((int(m.group(1), 16), cur, m.group(2)))
    return lines

def main():
    rep, kern, obj, src = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    hdr, data = sass_rows(rep, kern)
    ix = {h: i for i, h in enumerate(hdr)}
    table = line_table(obj, kern)
    base = int(data[0][ix["Address"]], 16)
    by_off = {off: ln for off, ln, _ in table}
    agg, stall = collections.Counter(), collections.defaultdict(collections.Counter)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = 0
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        n = float(r[ix["# Samples"]] or 0)
        ln = by_off.get(off)
        agg[ln] += n
        tot += n
        for c in stall_cols:
            stall[ln][c] += float(r[ix[c]] or 0)
    text = open(src).read().splitlines()
    print(f"total samples {tot:.0f}")
    for ln, n in agg.most_common(top):
        s = stall[ln].most_common(2)
        code = text[ln[1] - 1].strip()[:95] if ln and ln[0] == os.path.basename(src) else str(ln)
        print(f"{n:7.0f} {100 * n / tot:5.1f}%  L{ln[1] if ln else 0:<4d} {code:95s} {s[0][0][6:]}:{s[0][1]:.0f} {s[1][0][6:]}:{s[1][1]:.0f}")

if __name__ == "__main__":
    main()
