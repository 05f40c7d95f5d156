"""Per-source-line stall breakdown of one kernel out of an .ncu-rep (run here, no GPU needed).

  python tools/ncu_lines.py gpurun_out/step.ncu-rep path/to/lib.so '_ZN2as6k_stepILi0ELi0ELb1EEEvNS_8StepArgsE' [top] [name]

`name` (a substring of the demangled kernel name, default "k_step") picks the launch when the report holds several
kernels; the first launch that matches is analysed.

The report gives stall samples per SASS instruction; `nvdisasm -g` of the same cubin gives the source line of every
instruction (both list the kernel's instructions in address order).  Prints the kernel's headline metrics, the stall
reasons summed over the kernel, where the "no instruction" stalls sit (after a branch / at a reconvergence point /
elsewhere), the dynamic opcode mix and the hottest source lines.  K = as_step_kernel.cuh, M = as_math.cuh.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def line_map(lib, mangled):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = run(["nvdisasm", "-g", os.path.join(tmp, cubin)])
    lines, cur, on = [], None, False
    for l in txt.splitlines():
        if l.startswith(".text."):
            on = l.strip() == f".text.{mangled}:"
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            f = m.group(1).split("/")[-1].replace("as_step_kernel.cuh", "K").replace("as_math.cuh", "M")
            cur = (f, int(m.group(2)))
        elif re.match(r"\s*/\*[0-9a-f]+\*/", l):
            lines.append(cur)
    return lines


def main():
    rep, lib, mangled = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    want = sys.argv[5] if len(sys.argv) > 5 else "k_step"
    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
    h, u = raw[0], raw[1]
    kcol = h.index("Kernel Name")
    r = next((x for x in raw[2:] if want in x[kcol]), raw[2])
    for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
              "smsp__average_warp_latency_per_inst_issued.ratio"):
        if n in h:
            print(f"{n:70s} {r[h.index(n)]:>24s} {u[h.index(n)]}")
    src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    # one section per launch: a "Kernel Name" row, the column header, then one row per instruction
    starts = [i for i, x in enumerate(src) if x and x[0] == "Kernel Name"]
    pick = next((i for i in starts if want in src[i][1]), starts[0])
    end = next((i for i in starts if i > pick), len(src))
    hdr, data = src[pick + 1], [x for x in src[pick + 2:end] if x]
    ix = {n: i for i, n in enumerate(hdr)}

    def g(row, n):
        try:
            return int(row[ix[n]])
        except (ValueError, KeyError, IndexError):
            return 0

    lines = line_map(lib, mangled)
    if len(lines) != len(data):
        print(f"warning: {len(lines)} instructions in the library, {len(data)} in the report; no line mapping")
        lines = [None] * len(data)
    names = ["stall_barrier", "stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb",
             "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_math", "stall_dispatch",
             "stall_lg", "stall_mio"]
    tot = sum(g(x, "# Samples") for x in data)
    print(f"--- stall samples: {tot}")
    for n in names:
        v = sum(g(x, n) for x in data)
        print(f"  {n:24s} {v:6d} {100.0 * v / max(tot, 1):5.1f}%")
    cls = collections.Counter()
    dyn = collections.Counter()
    per = collections.defaultdict(collections.Counter)
    for i, x in enumerate(data):
        op = re.sub(r"^(@!?U?P\d\s+)", "", x[1].strip()).split()[0].split(".")[0] if x[1].strip() else ""
        dyn[op] += g(x, "Instructions Executed")
        ni = g(x, "stall_no_inst")
        if ni:
            prev = data[i - 1][1].strip() if i else ""
            if re.match(r"BRA", prev):
                cls["after an unconditional branch"] += ni
            elif "BSYNC" in x[1]:
                cls["at a reconvergence point"] += ni
            else:
                cls["elsewhere"] += ni
        for n in names:
            per[lines[i]][n] += g(x, n)
        per[lines[i]]["tot"] += g(x, "# Samples")
        per[lines[i]]["inst"] += g(x, "Instructions Executed")
    print("--- 'no instruction' stalls:", dict(cls))
    ti = sum(dyn.values())
    print(f"--- dynamic warp instructions: {ti}")
    print("  " + "  ".join(f"{k} {100.0 * v / ti:.1f}%" for k, v in dyn.most_common(18)))
    print("--- hottest source lines (samples; instructions; main reasons)")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]["tot"])[:top]:
        why = " ".join(f"{n[6:]}={v[n]}" for n in names if v[n] * 8 > v["tot"] and v[n] > 5)
        print(f"  {str(k):16s} {v['tot']:6d} {100.0 * v['tot'] / max(tot, 1):5.1f}%  inst {100.0 * v['inst'] / ti:4.1f}%  {why}")


if __name__ == "__main__":
    main()
