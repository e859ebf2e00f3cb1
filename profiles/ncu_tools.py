"""Helpers to read an .ncu-rep here (no GPU needed):
  python profiles/ncu_tools.py summary <rep>            # key metrics per captured launch
  python profiles/ncu_tools.py hot <rep> <kernel-regex> [launch-index]   # top SASS lines by stall samples
"""
import csv
import io
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_elapsed.max',
        'l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum', 'lts__t_sectors_op_read.sum',
        'lts__t_sectors_op_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def summary(rep):
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("-----")
        for w in KEEP:
            if w in idx:
                print(f"{w:72s} {r[idx[w]][:70]:>24s} {units[idx[w]]}")


def hot(rep, kernel, launch=0, top=40):
    out = run(["-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}", "--launch-skip", str(launch),
               "--launch-count", "1"])
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    i_src, i_s, i_ex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    data, seen = [], set()
    for r in rows[2:]:
        if r[0] in seen:
            continue
        seen.add(r[0])
        try:
            data.append((int(r[i_s]), int(r[i_ex]), len(data), r[i_src].strip()))
        except ValueError:
            pass
    tot = sum(d[0] for d in data)
    print(f"# {rows[0][1][:100]}: {tot} stall samples over {len(data)} SASS lines, "
          f"{sum(d[1] for d in data)} warp instructions")
    for s, ex, k, src in sorted(data, reverse=True)[:top]:
        print(f"{s:7d} {100 * s / max(tot, 1):5.1f}%  ex={ex:9d}  #{k:5d}  {src[:100]}")


def lines(rep, kernel, launch=0, top=40):
    """Stall samples and executed warp instructions aggregated per CUDA source line (file:line)."""
    out = run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kernel}",
               "--launch-skip", str(launch), "--launch-count", "1"])
    agg, fname, hdr = {}, None, None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            hdr = r; i_s = hdr.index('Warp Stall Sampling (All Samples)'); i_ex = hdr.index('Instructions Executed'); continue
        if hdr is None or r[0] == "Function Name":
            continue
        try:
            key = (fname, int(r[0]))
            s, ex = int(r[i_s] or 0), int(r[i_ex] or 0)
        except (ValueError, IndexError):
            continue
        a = agg.setdefault(key, [0, 0, r[1].strip()])
        a[0] += s; a[1] += ex
    tot_s = sum(a[0] for a in agg.values()); tot_ex = sum(a[1] for a in agg.values())
    print(f"# {kernel}: {tot_s} stall samples, {tot_ex} warp instructions")
    for (f, ln), (s, ex, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{s:7d} {100 * s / max(tot_s, 1):5.1f}%  ex={ex:11d} {100 * ex / max(tot_ex, 1):5.1f}%  {f}:{ln}  {src[:90]}")


if __name__ == "__main__":
    if sys.argv[1] == "summary":
        summary(sys.argv[2])
    elif sys.argv[1] == "lines":
        lines(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        hot(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
