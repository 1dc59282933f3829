"""Turn the raw ncu dump of one bench frame (tools/final_profile.sh) into profiles/<round>_traffic.json (DRAM bytes per
launch, read by bench.py) and a markdown table (stdout)."""
import csv, json, sys
raw, out_json = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
def g(r, k, f=float):
    try:
        return f(r[col[k]].replace(",", ""))
    except Exception:
        return None
names = {"morton_hist_kernel": "morton_hist_kernel", "lsd_sort_kernel": "lsd_sort_kernel", "transform_kernel": "transform_kernel",
         "collide_kernel": "collide_kernel"}
traffic = {}
print("| kernel | grid x block | time (us) | DRAM rd (MB) | DRAM wr (MB) | regs | occ % | DRAM % | L2 % | L1 % | warp-inst | stalls/issue: long-sb, barrier, short-sb, mio, lg, wait |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    kn = r[col["Kernel Name"]]
    key = None
    if "tree_emit_kernel" in kn:
        key = "tree_emit_kernel<build>" if ("tree_emit_kernel<1>" in kn or "<(bool)1>" in kn or "<true>" in kn) else "tree_emit_kernel<refit>"
    else:
        for k in names:
            if k in kn:
                key = k
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    unit = rows[1][col["dram__bytes_read.sum"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    unit_w = rows[1][col["dram__bytes_write.sum"]]
    scale_w = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit_w, 1)
    t = g(r, "gpu__time_duration.sum")
    tunit = rows[1][col["gpu__time_duration.sum"]]
    t_us = t * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(tunit, 1)
    if key and rd is not None:
        traffic.setdefault(key, []).append(int(rd * scale + (wr or 0) * scale_w))
    st = lambda n: g(r, f"smsp__average_warps_issue_stalled_{n}_per_issue_active.ratio")
    print(f"| {kn[:60]} | {r[col['Grid Size']]} x {r[col['Block Size']]} | {t_us:.1f} | {rd*scale/1e6:.1f} | {(wr or 0)*scale_w/1e6:.1f} | "
          f"{g(r,'launch__registers_per_thread',int)} | {g(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{g(r,'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g(r,'smsp__inst_executed.sum',float):.0f} | "
          f"{st('long_scoreboard'):.1f}, {st('barrier'):.1f}, {st('short_scoreboard'):.1f}, {st('mio_throttle'):.1f}, {st('lg_throttle'):.1f}, {st('wait'):.1f} |")
json.dump({k: int(sum(v) / len(v)) for k, v in traffic.items()} |
          {"source": "ncu --set full --clock-control none, tools/profile_frame.py via tools/final_profile.sh: dram__bytes_read.sum + "
                     "dram__bytes_write.sum per launch, mean over the launches of the kernel in one frame"}, open(out_json, "w"), indent=1)
