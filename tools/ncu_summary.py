#!/usr/bin/env python
"""Summaries of an .ncu-rep for profiles/ (run where `ncu` is installed; no GPU needed):
   ncu_summary.py raw REPORT        key metrics of every captured launch
   ncu_summary.py lines REPORT [N]  the N source lines with the most executed instructions, per kernel (needs -lineinfo + --import-source)
   ncu_summary.py traffic REPORT KERNEL_REGEX STREAMS SHA16   JSON for bench.py's roofline.traffic (DRAM bytes per launch and stream)"""
import csv, io, json, re, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def cmd_raw(rep):
    hdr, units, rows = raw_rows(rep)
    for r in rows:
        name = r[hdr.index("Kernel Name")]
        print(f"== {name}")
        vals = {}
        for i, h in enumerate(hdr):
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(r[i] or 0) >= 0.3):
                print(f"  {h:95s} {r[i]:>16s} {units[i]}")
                vals[h] = (r[i], units[i])
        rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
        t = vals.get("gpu__time_duration.sum")
        if rd and wr and t:
            b = to_bytes(*rd) + to_bytes(*wr)
            ms = float(t[0]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(t[1], 1)
            print(f"  => DRAM bytes (read+write) {b / 1e9:.3f} GB, / duration {b / (ms / 1e3) / 1e9:.1f} GB/s")
        print()


def cmd_lines(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    per, cur_file, kern = {}, None, None
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) == 2 and r[0] == "Function Name":
            kern = r[1]
        elif len(r) >= 8 and r[0] not in ("", "Line No") and r[2] == "-":
            try:
                n, smp = int(r[7]), int(r[6])
            except ValueError:
                continue
            a = per.setdefault(kern, {}).setdefault((cur_file, int(r[0]), r[1].strip()[:100]), [0, 0])
            a[0] += n
            a[1] += smp
    for k, d in per.items():
        tot, stot = sum(v[0] for v in d.values()) or 1, sum(v[1] for v in d.values()) or 1
        print(f"== {k}: {tot} warp instructions executed, {stot} stall samples")
        for key, v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"  {v[0] / tot * 100:5.1f}% inst {v[1] / stot * 100:5.1f}% samples  {key[0]}:{key[1]}  {key[2]}")
        print()


def cmd_traffic(rep, regex, streams, sha):
    hdr, units, rows = raw_rows(rep)
    tot, n, names = 0.0, 0, []
    for r in rows:
        name = r[hdr.index("Kernel Name")]
        if not re.search(regex, name):
            continue
        i, j = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        tot += to_bytes(r[i], units[i]) + to_bytes(r[j], units[j])
        names.append(name)
        n += 1
    print(json.dumps({"kernel_source_sha16": sha, "streams": int(streams), "launches_summed": names,
                      "dram_bytes_per_launch_per_stream": tot / int(streams),
                      "source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of {n} launch(es) at {streams} streams ({rep.split('/')[-1]})"}, indent=1))


if __name__ == "__main__":
    c = sys.argv[1]
    if c == "raw":
        cmd_raw(sys.argv[2])
    elif c == "lines":
        cmd_lines(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
    else:
        cmd_traffic(*sys.argv[2:6])
