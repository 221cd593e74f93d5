#!/usr/bin/env python
"""Per-instruction view of one kernel of an ncu report: stall-reason totals, opcode mix,
hottest SASS instructions.   python tools/ncu_source.py <report.ncu-rep> <kernel-regex> [top]"""
import collections, csv, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None and row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] and len(row) >= len(cur["hdr"]) - 2:
        cur["rows"].append(row)
b = blocks[0]
h = b["hdr"]; ix = {k: i for i, k in enumerate(h)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
rows = b["rows"]
tot_s = sum(f(r, "# Samples") for r in rows); tot_i = sum(f(r, "Instructions Executed") for r in rows)
print(f"# {b['name']}: {len(rows)} SASS instrs, {tot_i:.0f} warp-insts executed, {tot_s:.0f} samples")
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(f(r, k) for r in rows) for k in stalls}
print("stalls:", ", ".join(f"{k[6:]}={v / max(tot_s, 1):.1%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot_s))
mix = collections.Counter()
for r in rows:
    op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0] if op else "?"
    mix[op] += f(r, "Instructions Executed")
print("opcode mix:", ", ".join(f"{k}={v / tot_i:.1%}" for k, v in mix.most_common(18)))
print("shared: wavefronts", sum(f(r, "L1 Wavefronts Shared") for r in rows), "ideal", sum(f(r, "L1 Wavefronts Shared Ideal") for r in rows))
print("global sectors:", sum(f(r, "L2 Theoretical Sectors Global") for r in rows), "ideal", sum(f(r, "L2 Theoretical Sectors Global Ideal") for r in rows))
print(f"--- top {top} by samples")
for i, r in sorted(enumerate(rows), key=lambda ir: -f(ir[1], "# Samples"))[:top]:
    st = sorted(((f(r, k), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{i:5d} {f(r, '# Samples') / max(tot_s, 1):6.1%} exec={f(r, 'Instructions Executed'):9.0f}  {r[ix['Source']].strip()[:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
