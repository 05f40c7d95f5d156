"""Static SASS opcode summary of the built library (run here, no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_opcodes.txt
Shows which kernels use the TMA bulk-copy engine (UBLKCP), mbarriers (SYNCS), cp.async (LDGSTS), the L2::64B fill-size
qualifier (LTC64B) -- and that nothing uses tensor cores (no HMMA / UTC*MMA): nothing on the path is a contraction."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "allsteps_isaaclab_b200", "liballsteps_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
rows, tot = [], collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
    c = collections.Counter()
    for o in ops:
        base = o.split(".")[0]
        c[base] += 1
        if "LTC64B" in o:
            c["LTC64B"] += 1
        if base == "UBLKCP":
            c["UBLKCP.S.G" if ".S.G" in o else "UBLKCP.G.S"] += 1
        if "MMA" in base:
            c["any MMA"] += 1
    rows.append((name, len(ops), c))
    tot.update(c)
print("SASS opcode summary of allsteps_isaaclab_b200/liballsteps_b200.so (cuobjdump -sass, sm_100a; static counts)")
print("whole library: " + ", ".join(f"{k} {tot[k]}" for k in
      ["UBLKCP", "UBLKCP.S.G", "UBLKCP.G.S", "UBLKPF", "SYNCS", "LDGSTS", "LTC64B", "LDG", "STG", "LDS", "STS", "RED",
       "ATOM", "ATOMG", "REDUX", "MUFU", "BAR", "NANOSLEEP", "any MMA"]))
print("UBLKCP = cp.async.bulk (TMA 1-D bulk copy; S.G global->shared, G.S shared->global), UBLKPF = cp.async.bulk.prefetch.L2,")
print("SYNCS = mbarrier operations, LDGSTS = cp.async, LTC64B = loads with the L2::64B fill-size qualifier.\n")
print(f"{'kernel':96s} {'instr':>6s} {'TMA in/out':>10s} {'mbar':>5s} {'LTC64B':>6s} {'LDG':>4s} {'STG':>4s} {'LDS':>4s} {'STS':>4s} "
      f"{'RED+ATOM':>8s} {'MUFU':>5s} {'BAR':>4s}")
for name, n, c in sorted(rows, key=lambda r: -r[1]):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().replace("as::", "")
    print(f"{dem[:96]:96s} {n:6d} {c['UBLKCP.S.G']:6d}/{c['UBLKCP.G.S']:<3d} {c['SYNCS']:5d} {c['LTC64B']:6d} {c['LDG']:4d} {c['STG']:4d} "
          f"{c['LDS']:4d} {c['STS']:4d} {c['RED'] + c['ATOM'] + c['ATOMG']:8d} {c['MUFU']:5d} {c['BAR']:4d}")
