"""Attribute ncu per-SASS-instruction counts to CUDA source lines.

usage: sass_by_line.py <ncu --page source --csv dump> <push.cu> <kernel substring> [GPAT_STRICT]
Compiles the .cu to a cubin with -lineinfo, disassembles it with nvdisasm -g, pairs the n-th
SASS instruction of the kernel with the n-th row of the ncu dump (same binary, same order) and
prints executed warp-instructions and stall samples per source line.
"""
import collections, csv, re, subprocess, sys, os, tempfile

dump, cu, kname = sys.argv[1:4]
strict = sys.argv[4] if len(sys.argv) > 4 else "0"
tmp = tempfile.mkdtemp()
cubin = os.path.join(tmp, "k.cubin")
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                f"-DGPAT_STRICT={strict}", "-cubin", "-o", cubin, cu], check=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], check=True, capture_output=True, text=True).stdout
lines = dis.splitlines()
# locate the kernel's text section
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
locs = []  # one (file, line) per instruction
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = os.path.basename(m.group(1))
        ln = int(m.group(2))
        m2 = re.search(r'inlined at "([^"]+)", line (\d+)', l)
        cur = (f, ln)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        locs.append(cur)
rows = list(csv.reader(open(dump)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
print(f"sass instrs: disasm {len(locs)}, ncu {len(body)}", file=sys.stderr)
ix_exec = hdr.index("Instructions Executed")
ix_smp = hdr.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0, 0])
n = min(len(locs), len(body))
tot = 0
for k in range(n):
    e = int(body[k][ix_exec] or 0)
    s = int(body[k][ix_smp] or 0)
    a = agg[locs[k]]
    a[0] += e; a[1] += s; a[2] += 1
    tot += e
src = {}
for (f, ln) in agg:
    if f not in src:
        for cand in (os.path.join(os.path.dirname(cu), f),):
            if os.path.exists(cand):
                src[f] = open(cand).read().splitlines()
tots = sum(a[1] for a in agg.values())
print(f"total executed warp-instructions {tot}, samples {tots}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "60"))]:
    text = src.get(f, [""] * (ln + 1))[ln - 1].strip()[:90] if f in src and ln - 1 < len(src[f]) else ""
    print(f"{100 * a[0] / tot:6.2f}% exec {100 * a[1] / max(tots, 1):6.2f}% smp  n={a[2]:4d} {f}:{ln}  {text}")
