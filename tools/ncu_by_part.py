"""Per-kernel, per-part summary of an ncu --csv launch list (time, DRAM bytes, tensor-pipe activity).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,\
sm__ops_path_tensor_op_hmma_src_fp16_dst_fp32_realtime.sum,sm__inst_executed_pipe_tensor_subpipe_hmma.sum \
        --clock-control none --csv --log-file gpurun_out/x_launches.csv python tools/profile_pass.py 640
    python tools/ncu_by_part.py gpurun_out/x_launches.csv

Tensor-pipe columns.  tcgen05 kernels do NOT populate sm__pipe_tensor_subpipe_hmma_cycles_active_realtime (the
round-1 lists showed 0.0), and on this driver sm__pipe_tensor_cycles_active_realtime reads "n/a" and
sm__ops_path_tensor_op_hmma_* reads 0 for UTCHMMA work (gpurun_out/r2d_launches_640.csv).  What IS counted is
sm__inst_executed_pipe_tensor_subpipe_hmma.sum = the number of tcgen05.mma instructions (UTCHMMA).  For the GEMM kernel
every instruction is one cta_group::2 MMA of 256 x BN x 16 that keeps BOTH SMs of the pair busy for BN/2 cycles
(4096 MAC/cycle/SM), so
    tensor busy % = instructions x BN / (SMs x elapsed cycles),   issued TFLOP/s = instructions x 2*256*BN*16 / time
(all three f16x3 passes and the padding columns included), with BN from the layer width (gemm_pick_block_n).  For the
attention kernel (cta_group::1, mixed N) only the instruction count is printed.
"""
import collections
import csv
import re
import sys


def short(n):
    return re.sub(r'\(.*', '', n).replace('void ', '').replace('pafuse::', '').replace('<unnamed>::', '')


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    hdr = rows[0]
    ik, im, iv, iid, iu = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Metric Unit'))
    L = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        d = L.setdefault(r[iid], {'k': short(r[ik])})
        try:
            v = float(r[iv].replace(',', ''))
        except ValueError:
            v = 0.0
        u = r[iu]
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
        d[r[im]] = v * scale
    part = -1
    names = ['body', 'face', 'hands']
    agg = collections.OrderedDict()
    for d in L.values():
        if d['k'].startswith('embed'):
            part += 1
        a = agg.setdefault((part, d['k']), [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get('gpu__time_duration.sum', 0.0)
        a[2] += d.get('dram__bytes_read.sum', 0.0)
        a[3] += d.get('dram__bytes_write.sum', 0.0)
        a[4] += d.get('sm__inst_executed_pipe_tensor_subpipe_hmma.sum', 0.0)
        a[5] += d.get('gpu__time_duration.sum', 0.0) * d.get('sm__cycles_elapsed.avg.per_second', 0.0) * 1e-6   # us * Hz -> cycles
    tot = sum(a[1] for a in agg.values())
    C = {'body': 384, 'face': 224, 'hands': 256}
    HDS = {'body': 48, 'face': 32, 'hands': 32}

    def block_n(N):
        for bn in range(256, 31, -32):
            if N % bn == 0:
                return bn
        return 0

    print(f"# {path}: {len(L)} launches, {tot / 1e3:.2f} ms summed (cold-cache, serialised under ncu)")
    print(f"{'part':6s}{'kernel':40s}{'n':>4s}{'avg us':>9s}{'share':>7s}{'rd MB':>9s}{'wr MB':>9s}{'DRAM GB/s':>10s}{'SM MHz':>8s}"
          f"{'UTCHMMA':>10s}{'tensor %':>9s}{'issued TF/s':>12s}")
    for (p, n), a in agg.items():
        nm = names[p % 3] if p >= 0 else '-'
        us = a[1] / a[0]
        mhz = a[5] / a[1] if a[1] else 0.0                      # cycles / us
        busy = tf = ''
        m = re.match(r'gemm_f16x3_kernel<(\d), (\d)', n)
        if m and nm in C and a[4] > 0:
            epi = int(m.group(1))
            N = {3: 24 * HDS[nm], 1: 2 * C[nm]}.get(epi, C[nm])
            bn = block_n(N)
            busy = f"{100 * a[4] * bn / (148 * a[5]):.1f}" if a[5] else ''
            tf = f"{a[4] * 2 * 256 * bn * 16 / a[1] / 1e6:.0f}"
        print(f"{nm:6s}{n:40s}{a[0]:4d}{us:9.1f}{100 * a[1] / tot:6.1f}%{a[2] / a[0] / 1e6:9.1f}{a[3] / a[0] / 1e6:9.1f}"
              f"{(a[2] + a[3]) / a[1] / 1e3:10.0f}{mhz:8.0f}{a[4] / a[0]:10.0f}{busy:>9s}{tf:>12s}")


if __name__ == '__main__':
    main(sys.argv[1])
