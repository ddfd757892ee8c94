"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept in profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_X.csv  > profiles/..._launches.md
  python tools/ncu_summary.py kernel   gpurun_out/prof_X.ncu-rep  > profiles/..._kernel.md
"""
import collections
import csv
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_tcgen05_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r["Metric Unit"], v)
        a = agg.setdefault(r["Kernel Name"].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print("\n(ncu --metrics gpu__time_duration.sum --clock-control none: serialised, cold-cache per-launch "
          "times -- compare shares, not absolutes.)")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(out.splitlines()))
    hdr = rd[0]
    print("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(rd) - 2)) + " |")
    print("|---|---|" + "---:|" * (len(rd) - 2))
    names = [r[hdr.index("Kernel Name")] for r in rd[2:]] if "Kernel Name" in hdr else []
    for w in RAW:
        if w in hdr:
            i = hdr.index(w)
            print("| %s | %s | %s |" % (w, rd[1][i], " | ".join(r[i] for r in rd[2:])))
    if names:
        print("\nkernels: " + "; ".join(sorted(set(n.split("(")[0] for n in names))))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
