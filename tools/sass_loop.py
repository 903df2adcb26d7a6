#!/usr/bin/env python3
"""Print the instruction mix of the hottest loop of one kernel in a cubin/.so.

usage: sass_loop.py <file> <kernel-name-substring> [marker-mnemonic] [--count=N] [-v]
The hot loop is taken as the innermost backward branch whose body contains the
most occurrences of `marker` (default VIMNMX3)."""
import re, subprocess, sys, collections
path, pat = sys.argv[1], sys.argv[2]
marker = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("-") else "VIMNMX3"
want_cnt = None          # --count N: pick the (smallest) loop with exactly N markers instead of the one with the most
for a in sys.argv:
    if a.startswith("--count="):
        want_cnt = int(a.split("=")[1])
txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = None
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat in name:
        body = f
        print("function:", name)
        break
if body is None:
    sys.exit("kernel not found")
ins = []
for line in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: k for k, (a, _) in enumerate(ins)}
best = None
for k, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_idx:
            lo = addr_idx[tgt]
            cnt = sum(1 for _, tt in ins[lo:k + 1] if re.search(r"\b%s\b" % re.escape(marker), tt.split()[0] if not tt.startswith("@") else tt.split()[1]))
            size = k + 1 - lo
            if want_cnt is not None and cnt != want_cnt:
                continue
            if cnt and (best is None or cnt > best[0] or (cnt == best[0] and size < best[1])):
                best = (cnt, size, lo, k)
if best is None:
    sys.exit("no loop with marker found")
cnt, size, lo, hi = best
print(f"loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {size} instructions, {cnt} x {marker}")
mix = collections.Counter()
for _, t in ins[lo:hi + 1]:
    parts = t.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    mix[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("IMAD", "VIADD", "VIMNMX")) and "." in op else "")] += 1
for op, c in mix.most_common():
    print(f"  {c:5d}  {op}   ({c / cnt:.2f}/cell)")
print(f"  total/cell: {size / cnt:.2f}")
if "-v" in sys.argv:
    for a, t in ins[lo:hi + 1]:
        print(f"{a:06x}  {t}")
