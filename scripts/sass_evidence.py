"""Regenerate profiles/<tag>_sass.md from the built library: mnemonic counts and short excerpts that prove the
tcgen05 / TMA / TMEM path (cuobjdump -sass; no GPU needed).   python scripts/sass_evidence.py r2"""
import re, subprocess, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
FN = "_ZN2mp15tc_chain_kernelILi16ELb1EEEvNS_9TcProgramENS_4TcIOE"
sass = subprocess.run(["cuobjdump", "-sass", "-fun", FN, "multiply_b200/libmultiply_b200.so"], capture_output=True,
                      text=True).stdout.splitlines()
ins = [l for l in sass if re.search(r"/\*[0-9a-f]{4,5}\*/", l) and ";" in l]
def mnem(l):
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    return m.group(1) if m else "?"
cnt = collections.Counter(mnem(l) for l in ins)
what = [("UTCHMMA", "tcgen05.mma kind::f16 (5th-gen tensor core MMA, operands from shared memory, accumulator in TMEM)"),
        ("UTCBAR", "tcgen05.commit -> mbarrier"), ("LDTM", "tcgen05.ld (TMEM -> registers)"),
        ("UBLKCP", "cp.async.bulk global -> shared (TMA engine, weight slots)"),
        ("UTCATOMSWS", "tcgen05.alloc / dealloc of TMEM columns"), ("SYNCS", "mbarrier arrive / try_wait"),
        ("CCTL", "discard.global.L2 (CCTL.E.RML2) and prefetch"),
        ("MUFU", "ex2 / lg2 / rcp of the softplus epilogue, sqrt of the normal"),
        ("STS", "operand stores into the swizzled A image (st.shared.v4, 32-bit addresses)"),
        ("ST", "generic stores (were the operand stores before they were written as st.shared)"),
        ("STL", "local-memory stores (tile prologue: dynamically indexed x[] / embedding arrays; none in the chunk loops)"),
        ("LDL", "local-memory loads (same)"), ("BAR", "named barriers (epilogue warps)"),
        ("F2FP", "fp32 -> fp16x2 packs of the hi/lo operand split"), ("STG", "sigma' / feature stash and outputs"),
        ("LDG", ""), ("LDS", ""), ("HMMA", "legacy mma.sync path (must be 0)")]
out = [f"# SASS evidence for `tc_chain_kernel<16,true>` ({tag} build)", "",
       f"`cuobjdump -sass multiply_b200/libmultiply_b200.so` (sm_100a), function `{FN}`; regenerate with "
       "`python scripts/sass_evidence.py`.", "", "| SASS mnemonic | count | what it is |", "|---|---:|---|"]
for k, w in what:
    out.append(f"| `{k}` | {cnt.get(k, 0)} | {w} |")
out += ["", f"Total instructions: {len(ins)}.", ""]
def excerpt(title, pat, before=2, after=2, which=0):
    idx = [i for i, l in enumerate(ins) if re.search(pat, l)]
    if not idx:
        return
    i = idx[min(which, len(idx) - 1)]
    out.extend([f"## {title}", "", "```"] + [l.split("/* 0x")[0].rstrip() for l in ins[max(0, i - before):i + after + 1]] + ["```", ""])
excerpt("MMA issue loop (one elected lane): descriptors in uniform registers, UTCHMMA per K-step", r"UTCHMMA", 2, 2, 1)
excerpt("commit to an mbarrier", r"UTCBAR", 1, 1)
excerpt("weight loader (bulk copy into the ring, completes on an mbarrier)", r"UBLKCP", 2, 2)
excerpt("epilogue: TMEM load", r"LDTM", 2, 2, 3)
excerpt("epilogue: operand image stores", r"STS\.128", 1, 2, 4)
excerpt("dead-scratch discard", r"CCTL\.E\.RML2", 2, 2)
open(f"profiles/{tag}_sass.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:30]))
