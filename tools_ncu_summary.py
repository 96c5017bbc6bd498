"""Condense an .ncu-rep (one kernel) into the numbers DESIGN.md / profiles/ quote: raw metrics, stall mix, opcode
mix with lane utilisation, and the hot source lines (needs the cubin of the profiled build for line info).
usage: python tools_ncu_summary.py report.ncu-rep [kernel-symbol-substring]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
sym = sys.argv[2] if len(sys.argv) > 2 else None


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("  %-66s %s %s" % (w, vals[i], units[i]))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
h = rows[1]
idx = {k: i for i, k in enumerate(h)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[idx[k]])
    except Exception:
        return 0.0


ti = sum(f(r, "Instructions Executed") for r in data)
tt = sum(f(r, "Thread Instructions Executed") for r in data)
print("static SASS instructions %d, executed %.3g warp-instr, %.2f active lanes/instr" % (len(data), ti, tt / max(ti, 1)))
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tot = {s: sum(f(r, s) for r in data) for s in stalls}
S = sum(tot.values())
print("stall mix:", ", ".join("%s %.1f%%" % (s[6:], 100 * v / S) for s, v in sorted(tot.items(), key=lambda x: -x[1])[:7]))
op, opt = collections.Counter(), collections.Counter()
for r in data:
    src = r[idx["Source"]].split()
    if not src:
        continue
    o = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    op[o] += f(r, "Instructions Executed")
    opt[o] += f(r, "Thread Instructions Executed")
print("opcode mix:", ", ".join("%s %.1f%% (%.0f lanes)" % (o, 100 * v / ti, opt[o] / max(v, 1)) for o, v in op.most_common(12)))
ex = sorted((f(r, "Instructions Executed") for r in data), reverse=True)
cum = 0
for i, v in enumerate(ex):
    cum += v
    if cum >= 0.9 * ti:
        print("hot set: top %d static instructions (%.0f KB) cover 90%% of executed instructions" % (i + 1, (i + 1) * 16 / 1024))
        break

if sym:
    subprocess.run("mkdir -p /tmp/cub && cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all /root/repo/source_b200/libraysect_b200.so >/dev/null 2>&1", shell=True)
    dis = subprocess.run("nvdisasm -g -c /tmp/cub/raysect_b200.sm_100a.cubin", shell=True, capture_output=True, text=True).stdout.splitlines()
    start = [i for i, l in enumerate(dis) if l.startswith(".text.") and sym in l]
    if start:
        loc, cur = {}, None
        for l in dis[start[0] + 1:]:
            if l.startswith(".text."):
                break
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/", l)
            if m:
                loc[int(m.group(1), 16)] = cur
        base = int(data[0][idx["Address"]], 16)
        agg = collections.defaultdict(lambda: [0, 0, 0, 0])
        for r in data:
            k = loc.get(int(r[idx["Address"]], 16) - base)
            a = agg[k]
            a[0] += f(r, "Instructions Executed"); a[1] += f(r, "Thread Instructions Executed"); a[2] += f(r, "# Samples"); a[3] += 1
        ts = sum(a[2] for a in agg.values())
        print("source lines by executed warp-instructions (lane fill tells where divergence costs):")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:40]:
            print("   %-16s:%-4s static %4d  inst %5.2f%%  samples %5.2f%%  lanes %.1f" % (k[0] if k else "?", k[1] if k else "", a[3], 100 * a[0] / ti, 100 * a[2] / max(ts, 1), a[1] / max(a[0], 1)))
        print("hot source lines (by stall samples):")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][2])[:22]:
            print("   %-16s:%-4s static %4d  inst %5.2f%%  samples %5.2f%%  lanes %.1f" % (k[0] if k else "?", k[1] if k else "", a[3], 100 * a[0] / ti, 100 * a[2] / max(ts, 1), a[1] / max(a[0], 1)))
