"""Development aid: aggregate an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --print-source sass)
by the noinline device functions of the world kernel, using the function offsets in the object's symbol table."""
import csv, re, subprocess, sys
src_csv, obj = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "world_kernelILi32E"
elf = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
funcs = []
for line in elf.splitlines():
    mm = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$_ZN3myo12" + kern + r"\S*\$_ZN3myo\d+(\w+?)(ILi|E)", line)
    if mm:
        funcs.append((int(mm.group(1), 16), int(mm.group(2), 16), mm.group(3)))
funcs.sort()
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iexe, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall_cols = {h: k for k, h in enumerate(hdr) if h.startswith("stall_")}
base = int(rows[2][ia], 16)
agg = {}
for r in rows[2:]:
    off = int(r[ia], 16) - base
    name = "<kernel body>"
    for o, sz, n in funcs:
        if o <= off < o + sz:
            name = n
    a = agg.setdefault(name, dict(samples=0, inst=0, thr=0, n=0, stalls={}))
    a["samples"] += int(r[isamp]); a["inst"] += int(r[iexe]); a["thr"] += int(r[ithr]); a["n"] += 1
    for h, k in stall_cols.items():
        if k < len(r) and r[k].isdigit():
            a["stalls"][h] = a["stalls"].get(h, 0) + int(r[k])
ts, ti = sum(a["samples"] for a in agg.values()), sum(a["inst"] for a in agg.values())
print(f"{'function':24s} {'sass':>6s} {'samples%':>8s} {'inst%':>7s} {'lanes':>6s}  top stalls")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    top = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:4]
    tops = " ".join(f"{h[6:]}:{100*v/max(a['samples'],1):.0f}%" for h, v in top)
    print(f"{n:24s} {a['n']:6d} {100*a['samples']/ts:8.1f} {100*a['inst']/ti:7.1f} {a['thr']/max(a['inst'],1):6.1f}  {tops}")
