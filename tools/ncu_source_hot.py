"""Hottest SASS lines of one kernel of an .ncu-rep (warp-stall samples with the dominant stall reason).

    python tools/ncu_source_hot.py gpurun_out/x.ncu-rep <kernel regex> [launch index] [top N]
"""
import csv
import io
import subprocess
import sys


def main():
    path, rx = sys.argv[1], sys.argv[2]
    idx = sys.argv[3] if len(sys.argv) > 3 else "1"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-id", f"::regex:{rx}:{idx}"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
    if not starts:
        raise SystemExit("no kernel matched")
    end = starts[1] if len(starts) > 1 else len(lines)
    print(lines[starts[0]][:160])
    rows = list(csv.reader(io.StringIO("\n".join(lines[starts[0] + 1:end]))))
    hdr = rows[0]
    i_src, i_all = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    i_exec = hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    tot = 0
    reason_tot = {}
    for n, r in enumerate(rows[1:]):
        try:
            s = int(r[i_all])
        except (ValueError, IndexError):
            continue
        tot += s
        reasons = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
        for i, h in stall_cols:
            reason_tot[h] = reason_tot.get(h, 0) + int(r[i] or 0)
        data.append((s, n, r[i_src].strip(), int(r[i_exec] or 0), reasons))
    print(f"total samples {tot}; by reason: " + ", ".join(f"{h[6:]} {100 * v / max(tot, 1):.1f}%" for h, v in
                                                         sorted(reason_tot.items(), key=lambda kv: -kv[1])[:8]))
    for s, n, src, ex, reasons in sorted(data, reverse=True)[:top]:
        rs = " ".join(f"{h[6:]}={v}" for v, h in reasons if v)
        print(f"{100 * s / max(tot, 1):5.1f}%  #{n:5d} exec {ex:9d}  {src[:70]:70s} {rs}")


if __name__ == "__main__":
    main()
