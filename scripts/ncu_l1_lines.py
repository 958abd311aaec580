"""Shared-memory wavefronts / global tag requests per CUDA source line and opcode of one kernel launch.
usage: ncu_l1_lines.py <report.ncu-rep> <cubin> <mangled-substring> <demangled-substring> <ctas> [launch#]"""
import csv, re, subprocess, sys
rep, cubin, mangled, pretty, ctas = sys.argv[1:6]
ctas = float(ctas); which = int(sys.argv[6]) if len(sys.argv) > 6 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
blocks, c = [], None
for line in raw:
    if line.startswith('"Kernel Name"'):
        c = [line]; blocks.append(c)
    elif c is not None:
        c.append(line)
blk = [b for b in blocks if pretty in b[0]][which]
rows = list(csv.reader(blk[1:])); hdr = rows[0]; body = rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], 0, False
for ln in dis:
    if ln.startswith(".text."): on = mangled in ln
    elif ln.startswith(".section"): on = False
    if not on: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s+(/\*[0-9a-f]+\*/\s+)?(@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]+", ln) and ";" in ln: lines.append(cur)
g = lambda r, k: int(r[ix[k]] or 0)
W = sum(g(r, "L1 Wavefronts Shared") for r in body); WI = sum(g(r, "L1 Wavefronts Shared Ideal") for r in body)
T = sum(g(r, "L1 Tag Requests Global") for r in body); E = sum(g(r, "Instructions Executed") for r in body)
print(f"per CTA: instr {E/ctas:.0f}  shared wavefronts {W/ctas:.0f} (ideal {WI/ctas:.0f})  global tag requests {T/ctas:.0f}")
agg = {}
for k, r in enumerate(body):
    src = r[ix["Source"]].strip()
    op = (src.split()[1] if src.startswith("@") else src.split()[0]) if src else "?"
    op = ".".join(op.split(".")[:2])
    key = (lines[k] if k < len(lines) else ("?", 0), op)
    a = agg.setdefault(key, [0, 0, 0, 0])
    a[0] += g(r, "Instructions Executed"); a[1] += g(r, "L1 Wavefronts Shared"); a[2] += g(r, "L1 Wavefronts Shared Ideal"); a[3] += g(r, "L1 Tag Requests Global")
for (ln, op), (e, w, wi, t) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][3]))[:24]:
    print(f"  {ln[0][:14]}:{ln[1]:<4} {op:10s} instr {e/ctas:7.1f} smem-wf {w/ctas:7.1f} (ideal {wi/ctas:7.1f}) gtag {t/ctas:7.1f}")
