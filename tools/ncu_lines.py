"""Joins an ncu SASS source page (per-instruction counts) with nvdisasm -g line info of the same build:
python tools/ncu_lines.py <sass.csv from `ncu -i rep --page source --csv --print-source sass`> <nvdisasm -g -c output> [kernel-substr]
Prints instructions executed and stall samples per source line, heaviest first."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ie = hdr.index('Instructions Executed'); sm = hdr.index('# Samples'); te = hdr.index('Thread Instructions Executed')
prof = [(r[1].strip(), int(r[ie]), int(r[sm]), int(r[te])) for r in data]
# disassembly: sequence of (file,line) per instruction for the wanted function
pat = sys.argv[3] if len(sys.argv) > 3 else None
cur = None; infn = False; lines = []
for l in open(sys.argv[2]):
    m = re.match(r'\s*\.section\s+\.text\.(\S+),', l)
    if m:
        infn = pat is None or pat in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: lines.append((cur, m.group(2).strip()))
print('profile instrs', len(prof), 'disasm instrs', len(lines), file=sys.stderr)
n = min(len(prof), len(lines))
mism = sum(1 for i in range(n) if prof[i][0].split()[0:1] != lines[i][1].replace('{','').split()[0:1])
print('opcode mismatches', mism, file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for i in range(n):
    a = agg[lines[i][0]]; a[0] += prof[i][1]; a[1] += prof[i][2]; a[2] += 1; a[3] += prof[i][3]
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print(f'total instr {tot/1e9:.2f}G samples {tots}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 45]:
    print(f'{str(k):40s} instr% {100*a[0]/tot:5.1f}  samples% {100*a[1]/tots:5.1f}  sass {a[2]:4d}  lanes {a[3]/max(1,a[0]):4.1f}')
