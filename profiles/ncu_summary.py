"""Text summary of an .ncu-rep for profiles/ (read here, no GPU): per captured launch the duration, DRAM bytes,
issue / pipe utilisation, occupancy, warp-stall reasons (pc sampling) and the hottest source lines.
  python profiles/ncu_summary.py gpurun_out/prof_C5_r02a.ncu-rep [kernel-regex] > profiles/r02_ncu_kn_C5.txt"""
import csv
import io
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_pipe_fma.sum", "sm__cycles_elapsed.max",
    "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "lts__t_sectors_op_red.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def raw_rows(rep):
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    return rows[0], rows[1], rows[2:]


def hot_lines(rep, kernel, skip, top=18):
    out = run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kernel}",
               "--launch-skip", str(skip), "--launch-count", "1"])
    agg, fname, hdr = {}, None, None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
            i_t = hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or r[0] == "Function Name":
            continue
        try:
            key = (fname, int(r[0]))
            s, ex, t = int(r[i_s] or 0), int(r[i_ex] or 0), int(r[i_t] or 0)
        except (ValueError, IndexError):
            continue
        a = agg.setdefault(key, [0, 0, 0, r[1].strip()])
        a[0] += s; a[1] += ex; a[2] += t
    ts, te = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    lines = [f"  source lines by warp-instructions executed ({te / 1e6:.1f} M warp instructions, {ts} stall samples):"]
    for (f, ln), (s, ex, t, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        lines.append(f"    {100 * ex / max(te, 1):5.1f}% inst  {100 * s / max(ts, 1):5.1f}% stall  lanes {t / max(ex, 1):4.1f}/32  "
                     f"{f}:{ln}  {src[:84]}")
    return lines


def main():
    rep = sys.argv[1]
    pattern = sys.argv[2] if len(sys.argv) > 2 else "."
    hdr, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    print(f"# {rep} (ncu --set full --clock-control none; cold-cache, serialised replays: counters, not timings)")
    for r in rows:
        name = r[idx["Kernel Name"]]
        if not re.search(pattern, name):
            continue
        short = re.sub(r"\(.*", "", name)
        k = seen.get(short, 0)
        seen[short] = k + 1
        print(f"\n== {name[:110]}  [launch {k} of this kernel]")
        for w in WANT:
            if w in idx and r[idx[w]] != "":
                print(f"  {w:66s} {r[idx[w]][:22]:>22s} {units[idx[w]]}")
        stalls = []
        for h, i in idx.items():
            m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+?)(_not_issued)?$", h)
            if m and not m.group(2):
                try:
                    stalls.append((float(r[i].replace(",", "")), m.group(1)))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls)
        if tot > 0:
            top = sorted(stalls, reverse=True)[:8]
            print("  warp stall reasons (pc samples): " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in top))
        if k == 0:
            base = re.sub(r"<.*", "", short.replace("void ", "")).split("::")[-1].strip()
            for ln in hot_lines(rep, base, 0):
                print(ln)


if __name__ == "__main__":
    main()
