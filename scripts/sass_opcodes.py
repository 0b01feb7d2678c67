"""SASS opcode histograms of the production kernels in libgpat_cuda.so (static counts, `cuobjdump -sass`).
usage: python scripts/sass_opcodes.py > profiles/<tag>_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stochastic_parker_b200", "csrc", "libgpat_cuda.so")
WANT = [  # (substring of the mangled name, label)
    ("push_kernel_coopILi5ELi0ELb0ELi3E", "push_kernel_coop<L3D, 0, 0, 3>  (C5: 3-D, record split at the 128-byte line, rolled rounds)"),
    ("push_kernel_coopILi4ELi0ELb0ELi7E", "push_kernel_coop<L2D, 0, 0, 7>  (C4: 2-D + momentum diffusion, side plane)"),
    ("push_kernel_coopILi0ELi0ELb0ELi7E", "push_kernel_coop<L2B, 0, 0, 7>  (C1, C2: the kernel bench.py times)"),
    ("push_kernel_coopILi0ELi0ELb0ELi5E", "push_kernel_coop<L2B, 0, 0, 5>  (C3: mag_dependency = 0)"),
    ("push_kernel_coopILi1ELi0ELb0ELi144E", "push_kernel_coop<L2E, 0, 0, alt_spec(FT, no maps)>  (focused transport 2-D, production build)"),
    ("push_kernel_coopILi0ELi0ELb0ELi240E", "push_kernel_coop<L2B, 0, 0, alt_spec(Parker, maps)>  (2-D Parker + turbulence maps)"),
    ("push_kernel_coopILi0ELi0ELb0ELi80E", "push_kernel_coop<L2B, 0, 0, alt_spec(1-D, no maps)>  (1-D)"),
    ("push_kernel_coopILi3ELi0ELb0ELi144E", "push_kernel_coop<L3E, 0, 0, alt_spec(FT, no maps)>  (focused transport 3-D, two CTAs per SM)"),
    ("pack_kernel", "pack_kernel  (gradients + record packing)"),
    ("tile_scatter_kernelILi0E", "tile_scatter_kernel<0>  (compaction)"),
    ("diag_kernel", "diag_kernel  (spectra + local histograms + counters)"),
]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = {}
for blk in re.split(r"\n\s*Function : ", txt)[1:]:
    name, body = blk.split("\n", 1)
    funcs[name.strip()] = body
print("SASS opcode histograms of the production kernels in libgpat_cuda.so (sm_100a), `cuobjdump -sass`, static counts "
      "(scripts/sass_opcodes.py).\n256-bit global loads (LDG.E.*.256), no-return FP64 reductions (RED/REDG ... F64), warp-match "
      "aggregation (MATCH.ANY), FP64 FMA pipeline (DFMA); no tensor-core opcodes (nothing on this path is a dense contraction).\n")
for key, label in WANT:
    hit = [n for n in funcs if key in n and "escaped" not in n]
    if not hit:
        print(f"== {label}\n   (not found: {key})\n")
        continue
    body = funcs[hit[0]]
    ops = collections.Counter()
    full = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body):
        full[m.group(1)] += 1
        ops[m.group(1).split(".")[0]] += 1
    special = {k: v for k, v in full.items() if re.match(r"(LDG|LDS|STS|RED|ATOM|MATCH|MUFU|F2F|REDUX|HMMA|UTC|UTMA|SHFL\.)", k)}
    print(f"== {label}\n   {hit[0][:110]}")
    print(f"   total {sum(ops.values())} instructions; top opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(18)))
    print("   memory / special: " + ", ".join(f"{k} x{v}" for k, v in sorted(special.items())) + "\n")
