"""Summarise ncu CSV outputs into the small text files kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/x_launches.csv   > profiles/x_launches.txt
    python tools/summarize_ncu.py kernel   gpurun_out/x.ncu-rep [regex] > profiles/x_kernel.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_xu.sum",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("pafuse::", "").replace("<unnamed>::", "")
    return name.strip()


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        k = short(r[ik])
        tot[k] += float(r[iv].replace(",", ""))
        cnt[k] += 1
    unit = "ns"
    total = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {total / 1e6:.3f} ms summed device time (cold-cache, serialised under ncu)")
    print(f"{'kernel':60s} {'launches':>9s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
    for k, v in tot.most_common():
        print(f"{k[:60]:60s} {cnt[k]:9d} {v / 1e6:10.3f} {v / cnt[k] / 1e3:9.1f} {100 * v / total:6.1f}%")


def kernel(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    print(f"# {path}: ncu --set full, one block per profiled launch")
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[ik]):
            continue
        print(f"\n## {short(r[ik])}  (id {r[0]})")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:75s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
