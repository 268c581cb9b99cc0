#!/usr/bin/env python
"""Digest an .ncu-rep into two small text files (key counters per kernel; per-source-line hot spots)
so that only kilobytes travel back from the GPU box.   usage: ncu_digest.py rep.ncu-rep out_prefix [frames]
Also writes <out_prefix>_k2_capture.json: the k2_scan counters bench.py's roofline block quotes (DRAM bytes, shared
wavefronts, bank-conflict replays, duration) tagged with the hash of csrc/ they were measured on; copy it to
profiles/k2_capture.json so that bench.py can tell whether the capture belongs to the tree it runs."""
import csv, json, os, subprocess, sys, io

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_read.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max']
STALLS = 'smsp__average_warps_issue_stalled_'


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(run(['ncu', '-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    with open(out + '_metrics.txt', 'w') as f:
        kn = hdr.index('Kernel Name')
        for r in rows[2:]:
            f.write('=== %s\n' % r[kn])
            for i, h in enumerate(hdr):
                if h in KEEP or (h.startswith(STALLS) and h.endswith('_per_issue_active.ratio')):
                    f.write('%-92s %s %s\n' % (h, r[i], units[i]))
    # machine-readable capture of the dominant kernel for bench.py
    def num(r, name):
        v = r[hdr.index(name)].replace(',', '')
        u = units[hdr.index(name)]
        mult = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0,
                'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0}.get(u, 1.0)
        return float(v) * mult
    k2 = [r for r in rows[2:] if 'k2_scan' in r[kn]]
    if k2:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from jda_b200 import buildinfo
        r = k2[0]
        cap = {'kernel': r[kn], 'source_sha': buildinfo.source_sha(), 'frames': int(sys.argv[3]) if len(sys.argv) > 3 else 256,
               'time_s': num(r, 'gpu__time_duration.sum'),
               'dram_bytes_read': num(r, 'dram__bytes_read.sum'), 'dram_bytes_write': num(r, 'dram__bytes_write.sum'),
               'smem_wavefronts': num(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
               'smem_bank_conflicts': num(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
               'warp_instructions': num(r, 'smsp__inst_executed.sum'),
               'lsu_data_pipe_pct': num(r, 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed'),
               'issue_active_pct': num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
               'how': 'ncu --set full --clock-control none, one k2_scan launch of `bench.py --batch <frames>` (tools/gpu_round.sh)'}
        with open(out + '_k2_capture.json', 'w') as f:
            json.dump(cap, f, indent=1)
    rows = list(csv.reader(io.StringIO(run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass']))))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == 'Function Name':
            cur = {'name': r[1], 'rows': [], 'hdr': None}; secs.append(cur)
        elif r and r[0] == 'Line No' and cur is not None and cur['hdr'] is None:
            cur['hdr'] = r
        elif cur is not None and cur['hdr'] is not None and len(r) >= len(cur['hdr']):
            cur['rows'].append(r)
    with open(out + '_hotspots.txt', 'w') as f:
        for sec in secs:
            h = sec['hdr']; il = h.index('Line No'); isamp = h.index('# Samples'); ii = h.index('Instructions Executed')
            ie = h.index('L1 Wavefronts Shared Excessive') if 'L1 Wavefronts Shared Excessive' in h else None
            iw = h.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in h else None
            agg = {}; ti = ts = 0
            for r in sec['rows']:
                try:
                    ln = int(r[il]); ins = int(r[ii] or 0); sm = int(r[isamp] or 0)
                except ValueError:
                    continue
                a = agg.setdefault(ln, [0, 0, r[1], 0, 0]); a[0] += ins; a[1] += sm; ti += ins; ts += sm
                if ie is not None:
                    try:
                        a[3] += int(r[ie] or 0); a[4] += int(r[iw] or 0)
                    except ValueError:
                        pass
            if ti < 100000:
                continue
            f.write('==== %s  warp-instructions %d  stall samples %d\n' % (sec['name'][:70], ti, ts))
            for ln, (ins, sm, src, ex, wv) in sorted(agg.items()):
                if ins > ti * 0.01 or sm > ts * 0.01:
                    f.write('%4d %5.1f%%inst %5.1f%%samp smem-wavefronts %7.1fM excess %6.1fM | %s\n' %
                            (ln, 100 * ins / ti, 100 * sm / max(ts, 1), wv / 1e6, ex / 1e6, src.strip()[:90]))


if __name__ == '__main__':
    main()
