"""Join an ncu source-page CSV (per SASS instruction: samples, stall reasons) with the line table of the
cubin (nvdisasm --print-line-info) and print the warp-stall samples per SOURCE LINE of a kernel.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel substring> [launch index] [top N]
The .so must be the one the report was captured from (same SASS offsets)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "html5-canvas-raytracer_b200", "csrc", "librtb200.so")


def line_table(kernel_sub):
    """{offset: (file line, sass text)} of the first function whose mangled name contains kernel_sub."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-c", "--print-line-info", os.path.join(tmp, f)], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        cur, line, inl = None, 0, False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur = m.group(1)
                continue
            if cur is None or kernel_sub not in cur:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.setdefault(cur, {})[int(m.group(1), 16)] = (line, m.group(2).strip())
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    sel = [k for k in kernels if ksub in k["name"]]
    k = sel[launch]
    hdr = k["rows"][0]
    data = k["rows"][1:]
    ci = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    mangled = "wf_trace_kernelILb1" if "(bool)1" in k["name"] else ("wf_trace_kernelILb0" if "(bool)0" in k["name"] else ksub)
    tabs = line_table(mangled)
    fn = max(tabs.values(), key=len) if tabs else {}
    base = int(data[0][ci["Address"]], 16)
    per_line = collections.Counter()
    per_line_stall = collections.defaultdict(collections.Counter)
    per_line_inst = collections.Counter()
    total = 0
    for r in data:
        off = int(r[ci["Address"]], 16) - base
        s = int(r[ci["# Samples"]] or 0)
        line = fn.get(off, (("?", 0), ""))[0]
        per_line[line] += s
        per_line_inst[line] += int(r[ci["Instructions Executed"]] or 0)
        total += s
        for h in stall_cols:
            v = int(r[ci[h]] or 0)
            if v:
                per_line_stall[line][h[6:]] += v
    print(f"{k['name'][:90]}  launch {launch}: {total} samples, {len(data)} SASS instructions")
    src = {}
    for line, s in per_line.most_common(top):
        f, n = line
        if f not in src:
            p = os.path.join(ROOT, "html5-canvas-raytracer_b200", "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][n - 1].strip()[:70] if 0 < n <= len(src[f]) else ""
        st = ", ".join(f"{a} {100 * b / max(s, 1):.0f}%" for a, b in per_line_stall[line].most_common(3))
        print(f"{100 * s / total:5.1f}%  {f}:{n:<5d} inst {per_line_inst[line]:>11d}  [{st}]  {text}")


if __name__ == "__main__":
    main()
