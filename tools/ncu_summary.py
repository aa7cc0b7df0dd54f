"""Key metrics of every launch in an .ncu-rep (from `ncu --set full ...`), as a small table for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--stalls] > profiles/rNN_x_ncu_summary.txt

--stalls adds the top warp-stall sites of the first launch (SASS instruction, share of samples, dominant reason).
--traffic-json OUT --algorithmic BYTES [--launch J] writes the DRAM traffic of launch J (default 0) as the small JSON
bench.py reads for `roofline.traffic` (profiles/r02_spconv_hl_ncu_traffic.json).
"""
import csv
import io
import json
import subprocess
import sys

UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue_active_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, body = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    print(f"# {rep}: {len(body)} launches (ncu --set full --clock-control none; per-launch, cold-cache, serialised)")
    for j, r in enumerate(body):
        name = r[kn].replace("void ", "").replace("<unnamed>::", "")
        print(f"launch {j}: {name[:name.index('(') if '(' in name else 60]}")
        for m, label in METRICS:
            if m in hdr:
                i = hdr.index(m)
                v = r[i]
                try:
                    v = f"{float(v):.3f}".rstrip("0").rstrip(".")
                except ValueError:
                    pass
                print(f"    {label:16s} {v} {units[i]}")
    if "--traffic-json" in sys.argv:
        out = sys.argv[sys.argv.index("--traffic-json") + 1]
        j = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
        alg = float(sys.argv[sys.argv.index("--algorithmic") + 1]) if "--algorithmic" in sys.argv else None
        r = body[j]

        def nbytes(metric):
            i = hdr.index(metric)
            return float(r[i]) * UNIT_BYTES[units[i]]
        name = r[kn].replace("void ", "").replace("<unnamed>::", "")
        i_d = hdr.index("gpu__time_duration.sum")
        with open(out, "w") as f:
            json.dump({"source": rep, "launch": f"launch {j}: {name[:name.index('(') if '(' in name else 60]}, grid {r[hdr.index('launch__grid_size')]}, "
                                                 f"{r[i_d]} {units[i_d]} under ncu",
                       "dram_bytes_read": nbytes("dram__bytes_read.sum"), "dram_bytes_write": nbytes("dram__bytes_write.sum"),
                       "algorithmic_bytes": alg}, f, indent=1)
            f.write("\n")
    if "--stalls" in sys.argv:
        src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--launch-skip", "0", "--launch-count", "1"]))))
        h = src[1]
        i_s, i_src = h.index("# Samples"), h.index("Source")
        cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        b = [r for r in src[2:] if len(r) > i_s and r[i_s].isdigit()]
        if len(b) > 200 and b[0][i_src] == b[len(b) // 2][i_src]:
            b = b[:len(b) // 2]        # the report lists the function twice
        tot = sum(int(r[i_s]) for r in b) or 1
        print(f"# top stall sites of launch 0 ({tot} samples)")
        for r in sorted(b, key=lambda r: -int(r[i_s]))[:12]:
            st = sorted(((h[c][6:], int(r[c])) for c in cols if r[c].isdigit() and int(r[c]) > 0), key=lambda kv: -kv[1])[:1]
            print(f"    {100 * int(r[i_s]) / tot:5.1f}%  {r[i_src].strip()[:56]:56s} {st[0][0] if st else ''}")


if __name__ == "__main__":
    main()
