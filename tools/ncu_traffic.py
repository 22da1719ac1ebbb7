"""development helper: per-kernel totals of an .ncu-rep (ncu --set full capture) — duration, DRAM / L2 / L1 bytes, hit rates,
issue activity, occupancy, lanes per instruction, top stalls — as text (for profiles/*.txt) and as the JSON entry bench.py
reads from profiles/r2_traffic.json. Runs here (no GPU needed): python tools/ncu_traffic.py REP [--group NAME --per N]"""
import argparse, collections, csv, io, json, re, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--group", default=None, help="also print one JSON entry summing ALL captured launches under this kernel label")
ap.add_argument("--per", type=int, default=0, help="number of logical launches the group stands for (e.g. bounces captured); default: launches of the first kernel")
args = ap.parse_args()
raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def val(r, name, unit_scale=True):
    if name not in idx or r[idx[name]] in ("", "n/a"):
        return None
    try:
        v = float(r[idx[name]].replace(",", ""))
    except ValueError:          # "no data" on launches too short to sample
        return None
    u = units[idx[name]]
    if unit_scale:
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1, "us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3,
                 "Gbyte/s": 1e9, "Tbyte/s": 1e12, "Mbyte/s": 1e6, "Kbyte/s": 1e3, "byte/s": 1}
        for k, s in scale.items():
            if u == k:
                return v * s
    return v


SUM = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "SM_B.TriageCompute.l1tex__t_sectors.sum", "smsp__inst_executed.sum"]
AVG = ["lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
per = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("kdpt::", "").strip()
    d = per.setdefault(name, {"launches": 0, **{k: 0.0 for k in SUM}, **{k: 0.0 for k in AVG}})
    d["launches"] += 1
    w = val(r, "gpu__time_duration.sum") or 0.0
    for k in SUM:
        d[k] += val(r, k) or 0.0
    for k in AVG:
        d[k] += (val(r, k, False) or 0.0) * w       # duration-weighted


def show(name, d):
    ms = d["gpu__time_duration.sum"]
    print(f"{name}: {d['launches']} launch(es), {ms:.3f} ms")
    print(f"  DRAM read {d['dram__bytes_read.sum'] / 1e6:.1f} MB, written {d['dram__bytes_write.sum'] / 1e6:.1f} MB -> "
          f"{(d['dram__bytes_read.sum'] + d['dram__bytes_write.sum']) / ms / 1e6:.1f} GB/s;  L2 {(32 * d['lts__t_sectors.sum']) / 1e6:.1f} MB -> {(32 * d['lts__t_sectors.sum']) / ms / 1e6:.1f} GB/s;"
          f"  L1 {(32 * d['SM_B.TriageCompute.l1tex__t_sectors.sum']) / 1e6:.1f} MB -> {(32 * d['SM_B.TriageCompute.l1tex__t_sectors.sum']) / ms / 1e6:.1f} GB/s")
    a = {k: d[k] / ms for k in AVG}
    print(f"  L1 hit {a[AVG[1]]:.1f} %, L2 hit {a[AVG[0]]:.1f} %, issue active {a[AVG[2]]:.1f} %, warps active {a[AVG[3]]:.1f} % of peak, "
          f"{a[AVG[4]]:.2f} of 32 lanes per instruction, {a[AVG[5]]:.0f} registers, {d['smsp__inst_executed.sum'] / 1e6:.1f} M warp instructions")
    print(f"  stalls (warps per issue): long scoreboard {a[AVG[6]]:.2f}, wait {a[AVG[7]]:.2f}, no instruction {a[AVG[8]]:.2f}, branch {a[AVG[9]]:.2f}, "
          f"short scoreboard {a[AVG[10]]:.2f}, barrier {a[AVG[11]]:.2f}, not selected {a[AVG[12]]:.2f}")
    return a


for name, d in per.items():
    show(name, d)
if args.group:
    tot = {"launches": 0, **{k: 0.0 for k in SUM}, **{k: 0.0 for k in AVG}}
    for d in per.values():
        for k in SUM + AVG:
            tot[k] += d[k]
    first = next(iter(per.values()))
    n = args.per or first["launches"]
    tot["launches"] = n
    print("---- all captured kernels together, per logical launch (%d):" % n)
    a = show(args.group, tot)
    ms = tot["gpu__time_duration.sum"]
    hbm = 6553.0; l2 = 18457.4
    dram = tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]
    print(json.dumps({"kernel": args.group, "dram_bytes_per_launch": dram / n, "ms_per_launch_under_ncu": ms / n,
                      "dram_gbs": dram / ms / 1e6, "dram_frac_of_peak": dram / ms / 1e6 / hbm,
                      "l2_gbs": (32 * tot["lts__t_sectors.sum"]) / ms / 1e6, "l2_frac_of_peak": (32 * tot["lts__t_sectors.sum"]) / ms / 1e6 / l2,
                      "l1_hit_pct": a[AVG[1]], "l2_hit_pct": a[AVG[0]], "issue_active_pct": a[AVG[2]], "occupancy_pct": a[AVG[3]],
                      "lanes_per_instruction": a[AVG[4]]}))
