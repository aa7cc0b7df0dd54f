"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`)
into per-kernel totals of ONE steady-state fragment.

    python tools/summarize_launches.py gpurun_out/launches.csv [fragment_index_from_end] > profiles/rNN_launches_summary.txt

A fragment is delimited by two consecutive `init_prune_kernel` launches (one per NeuConNet.forward).  ncu serialises
the launches and runs each one cold, so the per-launch times are NOT the bench's step time: compare SHARES.
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)                       # drop the argument list
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    m = re.match(r"([\w:]+?)(<.*)?$", name)
    base = m.group(1) if m else name
    if base.startswith("spconv_tc_kernel") or base.startswith("bp_fused_kernel"):
        return name[:60]                                     # keep the template arguments of the headline kernels
    return base[:70]


def main():
    path = sys.argv[1]
    back = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    header = None
    for r in rd:
        if header is None:
            if "Kernel Name" in r:
                header = {k: i for i, k in enumerate(r)}
            continue
        if len(r) < len(header):
            continue
        if r[header["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit = r[header["Metric Unit"]]
        val = float(r[header["Metric Value"]].replace(",", ""))
        scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1e-3)
        rows.append((r[header["Kernel Name"]], val * scale))
    marks = [i for i, (k, _) in enumerate(rows) if "init_prune_kernel" in k]
    if len(marks) <= 1:
        # `ncu --profile-from-start off`: bench.py brackets exactly one steady-state fragment with cudaProfilerStart/Stop
        lo, hi, seg = 0, len(rows), rows
        marks, back = [0], 0
    else:
        if len(marks) < back + 1:
            raise SystemExit(f"only {len(marks)} fragments in the list")
        lo, hi = marks[-back - 1], marks[-back]
        seg = rows[lo:hi]
    agg = collections.defaultdict(lambda: [0.0, 0])
    for k, us in seg:
        a = agg[short(k)]
        a[0] += us
        a[1] += 1
    tot = sum(a[0] for a in agg.values())
    print(f"# {path}: {len(rows)} launches listed, {len(marks)} fragments; fragment #{len(marks) - back} "
          f"(launches {lo}..{hi - 1}): {len(seg)} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised under ncu)")
    print("# us_total  share  launches  kernel")
    for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{us:10.1f} {100 * us / tot:6.1f}% {n:6d}  {k}")


if __name__ == "__main__":
    main()
