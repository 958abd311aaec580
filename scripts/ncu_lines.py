"""Executed warp-instructions / stall samples per CUDA source line of one kernel.

usage: ncu_lines.py <report.ncu-rep> <cubin> <mangled-kernel-substring> <demangled-kernel-substring> [launch#]
Joins the SASS page of the report (in address order) with `nvdisasm --print-line-info` of the cubin."""
import csv, re, subprocess, sys
rep, cubin, mangled, pretty = sys.argv[1:5]
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], 0, False
for ln in dis:
    if ln.startswith(".text."):
        on = mangled in ln
    elif ln.lstrip().startswith(".section") or ln.startswith(".section"):
        on = False if ".text." not in ln else (mangled in ln)
    if not on:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+(/\*[0-9a-f]+\*/\s+)?(@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]+", ln) and ";" in ln:
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
blocks, c = [], None
for line in raw:
    if line.startswith('"Kernel Name"'):
        c = [line]; blocks.append(c)
    elif c is not None:
        c.append(line)
blocks = [b for b in blocks if pretty in b[0]]
blk = blocks[which]
rows = list(csv.reader(blk[1:])); hdr = rows[0]
iEx, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
body = rows[1:]
print("sass rows", len(body), "disasm instrs", len(lines))
agg = {}
tot = totS = 0
for k, r in enumerate(body):
    key = lines[k] if k < len(lines) else ("?", 0)
    e, s = int(r[iEx]), int(r[iS])
    a = agg.setdefault(key, [0, 0, 0]); a[0] += e; a[1] += s; a[2] += 1
    tot += e; totS += s
src = {}
for key, (e, s, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    f, l = key
    if f not in src:
        try:
            src[f] = open(f"/root/repo/gcgcn_b200/csrc/{f}").read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][l - 1].strip()[:95] if 0 < l <= len(src[f]) else ""
    print(f"{100*e/tot:5.1f}% st{100*s/max(totS,1):5.1f}% n{cnt:4d} {f}:{l:<4} {text}")
