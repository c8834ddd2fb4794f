"""Per-source-line stall samples of one kernel: joins the SASS page of an .ncu-rep (`ncu --page source --csv`) with the
line table of the cubin (`nvdisasm -g`).  Runs on the CPU box.
    python tools/ncu_lines.py <report.ncu-rep> <object-with-kernel.o> <mangled-kernel-name> [top-N] [result-index]
(result-index: which launch of a multi-kernel report, 0-based in report order; default 0)"""
import csv, io, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# walk the function: "//## File "...", line N" markers followed by instructions "/*0010*/ ..."
lines = []  # (file, line, inlined-at chain) per instruction in order
infn = False
cur = ("?", 0)
for l in dis:
    if l.startswith("\t.section\t.text."):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# one section per profiled launch: "Kernel Name" line, header line ("Address", ...), then the instructions
sections, cur_rows = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur_rows = {"name": r[1] if len(r) > 1 else "", "hdr": None, "data": []}
        sections.append(cur_rows)
    elif cur_rows is not None and r and r[0] == "Address":
        cur_rows["hdr"] = r
    elif cur_rows is not None and r and r[0].startswith("0x"):
        cur_rows["data"].append(r)
sec = sections[which]
print("kernel:", sec["name"][:120])
hdr, data = sec["hdr"], sec["data"]
assert len(data) == len(lines), (len(data), len(lines))
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
tot = 0
for r, (f, ln, *_rest) in zip(data, lines):
    n = int(r[col["# Samples"]] or 0)
    tot += n
    a = agg.setdefault((f, ln), {"n": 0, "inst": 0, "st": {}})
    a["n"] += n
    a["inst"] += int(r[col["Instructions Executed"]] or 0)
    for h in stall_cols:
        v = int(r[col[h]] or 0)
        if v:
            a["st"][h[6:]] = a["st"].get(h[6:], 0) + v
srccache = {}
def src(f, ln):
    for root in ("vae_mdl_b200/csrc", "."):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in srccache:
                srccache[p] = open(p).read().splitlines()
            L = srccache[p]
            return L[ln - 1].strip()[:80] if 0 < ln <= len(L) else ""
    return ""
print(f"total samples {tot}, instructions {sum(a['inst'] for a in agg.values())}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
    st = ", ".join(f"{k} {v}" for k, v in sorted(a["st"].items(), key=lambda kv: -kv[1])[:3])
    print(f"{a['n']:7d} {100 * a['n'] / tot:5.1f}%  inst {a['inst']:9d}  {f}:{ln:<4d} {src(f, ln):80s} | {st}")
