"""Turns the ncu artefacts a GPU visit left under gpurun_out/ into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_bench.csv profiles/r01_launches_bench.txt
    python tools/summarize_ncu.py kernels  gpurun_out/r01_render.ncu-rep   profiles/r01_render_kernels.txt
    python tools/summarize_ncu.py traffic  gpurun_out/r02_k1_256.ncu-rep   profiles/r02_k1_256_traffic.json
        (DRAM bytes of the captured K1 launch + the hash of the kernel sources the .so was built from -- run it
         BEFORE touching the kernel sources again; bench.py refuses a capture whose hash is stale)
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]


def us(row):
    t = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return t / 1000 if u == "ns" else t * 1000 if u == "ms" else t * 1e6 if u == "s" else t


def launches(src, dst):
    rows = list(csv.DictReader(l for l in open(src) if not l.startswith("==")))
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += us(r)
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {len(rows)} launches, {total / 1e3:.2f} ms of "
                f"kernel time (cold-cache, serialised)\n# source: {src}\n")
        f.write(f"{'share':>7} {'total us':>11} {'n':>5} {'avg us':>9}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * v[1] / total:6.2f}% {v[1]:11.1f} {v[0]:5d} {v[1] / v[0]:9.1f}  {k[:150]}\n")


def kernels(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in ix]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, one line per captured launch; source: {src}\n")
        for c in cols:
            f.write(f"#   {c} [{units[ix[c]]}]\n")
        for r in rows[2:]:
            f.write(r[ix["Kernel Name"]].split("(")[0][-40:].ljust(42) + " ".join(f"{r[ix[c]]:>12.12}" for c in cols) + "\n")


def traffic(src, dst):
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    r = rows[2]  # first captured launch

    def nbytes(name):
        v, u = float(r[ix[name]].replace(",", "")), units[ix[name]].lower()
        return int(round(v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]))
    rec = {"kernel": r[ix["Kernel Name"]][:160], "nv": 3, "D": 256,
           "dram_bytes_read": nbytes("dram__bytes_read.sum"), "dram_bytes_write": nbytes("dram__bytes_write.sum"),
           "gpu_time_us_under_ncu": float(r[ix["gpu__time_duration.sum"]].replace(",", "")),
           "kernel_source_sha256_16": bench.k1_source_hash(), "kernel_sources": list(bench.K1_SOURCES),
           "source": f"{src} (ncu --set full --clock-control none, one launch)"}
    json.dump(rec, open(dst, "w"), indent=1)
    print(rec)


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
