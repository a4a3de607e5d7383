"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).

    python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches_summary.txt
    python scripts/summarize_ncu.py full     gpurun_out/prof_r1.ncu-rep profiles/r1_top_kernels_ncu_full.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    for k, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, k
            break
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    n = 0
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].split("(")[0]
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s, %d launches, total %.1f us\n" % (src, n, tot))
        f.write("%-64s %5s %12s %10s %7s\n" % ("kernel", "n", "total_us", "avg_us", "share"))
        for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-64s %5d %12.1f %10.1f %6.1f%%\n" % (name[:64], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; one block per captured launch\n# source: %s\n" % src)
        for r in data:
            f.write("---- %s   grid=%s block=%s\n" % (r[idx["Kernel Name"]][:90], r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
            for m in METRICS:
                if m in idx:
                    f.write("    %-72s %s %s\n" % (m, r[idx[m]], units[idx[m]]))
    print(open(dst).read()[:3000])


def entry_key(kernel_name):
    """C-ABI entry point / head count key used by bench.py's roofline table for an ncu kernel name, or None."""
    import re
    n = kernel_name
    m = re.search(r"aggregate_fwd_kernel<(?:\(int\))?(\d+)", n)
    if m:
        return "ngacf_aggregate_fwd/H" + m.group(1)
    m = re.search(r"stage_bwd_edges_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+)", n)
    if m:
        return "ngacf_stage_bwd_edges_%s/H%s" % ("users" if m.group(2) == "0" else "items", m.group(1))
    m = re.search(r"transform_tc_kernel<(?:\(int\))?(\d+), (?:\(int\))?0", n)
    if m:
        return "ngacf_transform_fwd/H" + m.group(1)
    m = re.search(r"transform_bwd_tc_kernel<(?:\(int\))?(\d+)", n)
    if m:
        return "ngacf_transform_bwd/H" + m.group(1)
    m = re.search(r"stage_bwd_prep_kernel<(?:\(int\))?(\d+)", n)
    if m:
        return "ngacf_stage_bwd_prep/H" + m.group(1)
    for k in ("score_pairs_bwd", "dropout_masks", "adam"):
        if k + "_kernel" in n:
            return {"adam": "ngacf_adam_step_dev"}.get(k, "ngacf_" + k)
    return None


def traffic(src, dst):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) per entry point -> json for bench.py's roofline.traffic"""
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.OrderedDict()
    for r in data:
        key = entry_key(r[idx["Kernel Name"]])
        if key is None:
            continue
        b = sum(float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += b
    res = {k: dict(dram_bytes_per_launch=v[1] / v[0], launches=v[0]) for k, v in agg.items()}
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
