"""Per-kernel instruction/stall share of the code regions between BAR.SYNC instructions (SASS order)."""
import csv, subprocess, sys
rep, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
blocks, cur = [], None
for line in raw:
    if line.startswith('"Kernel Name"'):
        cur = [line]; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
for blk in blocks:
    name = blk[0][15:110]
    if pat not in name:
        continue
    rows = list(csv.reader(blk[1:])); hdr = rows[0]
    iS, iSrc, iEx = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[iEx]) for r in rows[1:]); totS = sum(int(r[iS]) for r in rows[1:])
    print("=====", name, "warp-instr", tot, "samples", totS)
    seg = segS = nh = 0; k0 = 0; ops = {}
    for k, r in enumerate(rows[1:]):
        e, s = int(r[iEx]), int(r[iS]); seg += e; segS += s
        src = r[iSrc].strip()
        op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0] if src else "?"
        ops[op] = ops.get(op, 0) + e
        if "HMMA" in src: nh += e
        if "BAR.SYNC" in src or k == len(rows) - 2:
            top = ", ".join(f"{o}:{100*c/max(seg,1):.0f}%" for o, c in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
            print(f"  sass[{k0:5d}-{k:5d}] instr {100*seg/tot:5.1f}%  stalls {100*segS/totS:5.1f}%  hmma={nh:8d}  {top}")
            seg = segS = nh = 0; k0 = k + 1; ops = {}
