"""Executed warp instructions per source line of one launch in an .ncu-rep (ncu_lines.py sorted by instruction count, plus a
per-file total).  python tools/ncu_inst_by_line.py <report> <object> <mangled-kernel-name> <result-index> [top]"""
import collections, contextlib, io, os, re, sys
rep, obj, kern, which = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
here = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(here, "ncu_lines.py")).read()
src = src.replace('key=lambda kv: -kv[1]["n"])[:top]', 'key=lambda kv: -kv[1]["inst"])[:top]')
sys.argv = ["ncu_lines.py", rep, obj, kern, "100000", which]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    exec(compile(src, "ncu_lines", "exec"))
lines = buf.getvalue().splitlines()
print("\n".join(lines[:2]))
tot = int(re.search(r"instructions (\d+)", lines[1]).group(1))
byfile, rows = collections.Counter(), []
for l in lines[2:]:
    m = re.match(r"\s*(\d+)\s+[\d.]+%\s+inst\s+(\d+)\s+(\S+):(\d+)\s+(.*?)\s+\|", l)
    if m:
        rows.append((int(m.group(2)), m.group(3), int(m.group(4)), m.group(5)[:80]))
        byfile[m.group(3)] += int(m.group(2))
print({k: f"{100 * v / tot:.1f}%" for k, v in byfile.most_common(10)})
for inst, f, ln, txt in rows[:top]:
    print(f"{100 * inst / tot:5.2f}%  {f}:{ln:<4d} {txt}")
