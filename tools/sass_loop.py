#!/usr/bin/env python
"""Extract the main loop of a kernel from `cuobjdump -sass` and print it with a per-pipe instruction histogram.

    python tools/sass_loop.py <kernel-name-regex> > profiles/r02_sass_<name>.txt

The loop is taken as the smallest backward-branch region that contains at least four REDUX instructions (the search
step of every engine: two for the selection, two or more for the rescan)."""
import collections, re, subprocess, sys
LIB = "neural-astar_b200/lib/libnastar_b200.so"
pat = re.compile(sys.argv[1])
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
fn, cur, funcs = None, [], {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        if fn: funcs[fn] = cur
        fn, cur = m.group(1), []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and fn:
        cur.append((int(m.group(1), 16), m.group(2).strip()))
if fn: funcs[fn] = cur
name = next(n for n in funcs if pat.search(subprocess.run(["c++filt", n], capture_output=True, text=True).stdout))
ins = funcs[name]
addr_ix = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)? (?:!?U?P\d, )?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr_ix:
        j = addr_ix[int(m.group(1), 16)]
        n_red = sum("REDUX" in x for _, x in ins[j:i + 1])
        if n_red >= 4 and (best is None or i - j < best[2] - best[1]):
            best = (n_red, j, i)
_, j, i = best
body = ins[j:i + 1]
ALU = ("IADD", "IADD3", "LOP", "LOP3", "SHF", "SHL", "SHR", "ISETP", "SEL", "PRMT", "LEA", "VIADD", "VIMNMX", "IMNMX", "FMNMX", "FSETP", "PLOP3", "IABS", "BMSK", "POPC", "FLO", "FSEL", "MOV", "CS2R", "VIADDMNMX", "P2R", "R2P", "SGXT")
FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FFMA2")
hist = collections.Counter()
for _, t in body:
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
    cls = ("alu" if op in ALU else "fma" if op in FMA else "lsu" if op in ("LDS", "STS", "LDG", "STG", "ATOMS", "LDSM", "LDC", "LDCU") else
           "redux" if "REDUX" in op else "shfl" if op in ("SHFL",) else "conv/sfu" if op in ("MUFU", "I2F", "F2I", "I2FP", "F2FP") else
           "branch" if op in ("BRA", "BSSY", "BSYNC", "WARPSYNC", "NOP", "BREAK", "CALL", "RET", "EXIT") else "uniform" if op.startswith("U") or op in ("S2UR", "R2UR", "VOTEU") else "other")
    hist[cls] += 1
print(f"# {subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()}")
print(f"# loop body: {len(body)} SASS instructions (static; predicated-off / not-taken paths included), addresses 0x{body[0][0]:x}..0x{body[-1][0]:x}")
print("# per-pipe histogram:", dict(hist.most_common()))
for a, t in body:
    print(f"/*{a:04x}*/  {t}")
