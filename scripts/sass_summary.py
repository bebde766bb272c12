"""SASS summary of the shipped library (no GPU needed): per kernel the instruction count and the counts of the mnemonics
that matter for the roofline discussion (256-/128-bit global loads, MUFU, conversions, shared-memory and global atomics,
cp.async, TMA / tensor-core mnemonics — expected absent), plus a short excerpt of the hot loops.
    python scripts/sass_summary.py > profiles/r2_sass_summary.md"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "piccolo_b200", "libpiccolo_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern = None; body = collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        kern = m.group(1); body[kern] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and kern:
        body[kern].append(m.group(2).strip())
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
WATCH = [("LDG.E.ENL2.256", r"LDG\.E\.ENL2\.256"), ("LDG.E.128", r"LDG\.E\.128"), ("LDG (all)", r"\bLDG"), ("LDS", r"\bLDS"), ("LDGSTS (cp.async)", r"LDGSTS"),
         ("MUFU", r"\bMUFU"), ("FFMA", r"\bFFMA\b"), ("HADD2.F32 (fp16->fp32)", r"HADD2\.F32"), ("PRMT", r"\bPRMT"), ("SHFL", r"\bSHFL"),
         ("ATOMS/ATOMG/RED", r"\b(ATOMS|ATOMG|ATOM|RED)\b"), ("BAR", r"\bBAR\b"), ("DADD/DFMA/DMUL", r"\b(DADD|DFMA|DMUL)\b"),
         ("TMA (UTMALDG/UBLKCP)", r"UTMALDG|UBLKCP"), ("tensor core (UTC*MMA/HMMA/LDTM)", r"UTC\w*MMA|HMMA|LDTM|IMMA")]
print("# SASS summary of `piccolo_b200/libpiccolo_b200.so` (sm_100a), `cuobjdump -sass`\n")
print("Kernels of this library only (CUB sort kernels omitted).  No TMA and no tensor-core mnemonic appears anywhere: the hot path is a "
      "gather + fp32 arithmetic, not a dense contraction (DESIGN.md 4).\n")
print("| kernel | instructions | " + " | ".join(n for n, _ in WATCH) + " |")
print("|---|---|" + "---|" * len(WATCH))
tot = collections.Counter()
for k, ins in body.items():
    if "pcl_" not in k:
        continue
    name = demangle(k)
    name = re.sub(r"\((int|bool)\)", "", name)
    name = re.sub(r">\(.*", ">", name) if ">(" in name else re.sub(r"\(.*", "", name)
    name = name.replace("void ", "")
    row = []
    for n, pat in WATCH:
        c = sum(1 for i in ins if re.search(pat, i)); row.append(c); tot[n] += c
    print(f"| `{name}` | {len(ins)} | " + " | ".join(str(c) for c in row) + " |")
print("\nTotals over the library's kernels: " + ", ".join(f"{n}: {tot[n]}" for n, _ in WATCH) + "\n")


def excerpt(pattern, title, first, n=28):
    for k, ins in body.items():
        if re.search(pattern, k):
            idx = next((i for i, s in enumerate(ins) if re.search(first, s)), None)
            if idx is None:
                continue
            print(f"## {title}\n\n`{demangle(k)[:140]}` — {n} instructions around the first `{first}`:\n\n```")
            for s in ins[max(0, idx - 8): idx + n - 8]:
                print("  " + s)
            print("```\n")
            return


excerpt(r"pcl_refine_persistent_kernelILi5ELi3E", "Fused refinement (F16D table, 3 candidates per block): texel fetch inside the evaluation loop", r"LDG\.E\.ENL2\.256")
excerpt(r"pcl_grid_score_kernelILi5E", "Structured-grid scoring (F16D table): member-rotation fetch", r"LDG\.E\.ENL2\.256")
excerpt(r"pcl_sample_kernelILi5ELb1E", "Generic fused forward+backward (F16D table)", r"LDG\.E\.ENL2\.256")
