"""Per-source-line view of an ncu capture: joins the SASS page of a report (ncu --page source --csv) with the line table of
the cubin (nvdisasm -g), in instruction order.  usage: python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]"""
import csv, io, re, subprocess, sys
from collections import defaultdict

def main(rep, cubin, kname, top=40):
    sass = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
    lines, cur, inside = [], None, False
    for ln in sass:
        if ln.startswith('.text.') or ln.strip().startswith('.section'):
            inside = kname in ln and '.text.' in ln
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if inside and re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
            lines.append(cur)
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = rows[1], rows[2:]
    ia, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
    if len(data) != len(lines):
        print('warning: %d SASS rows in the report, %d in the cubin' % (len(data), len(lines)))
    agg = defaultdict(lambda: [0, 0])
    for r, l in zip(data, lines):
        agg[l][0] += int(r[ia]); agg[l][1] += int(r[isamp])
    ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
    print('total warp instructions %d, samples %d' % (ti, ts))
    for l, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%-28s inst %5.2f%%  samples %5.2f%%' % ('%s:%d' % l if l else '?', 100.0 * v[0] / ti, 100.0 * v[1] / ts))
    return agg

if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
