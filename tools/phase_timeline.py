"""Summarise the [l3d] marks printed under L3D_DEBUG_PHASES=1 (last run in the log): per 7-frame chunk
the front / aggregation / back intervals in ms relative to the run's first enqueue."""
import re, sys
lines = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith('[l3d]')]
idx = len(lines) - 1
while idx >= 0 and 'mark' in lines[idx]: idx -= 1
while idx >= 0 and 'aggregation' in lines[idx]: idx -= 1
run = lines[idx + 1:]
ag = [tuple(map(float, re.findall(r'start ([\d.]+) ms, end ([\d.]+) ms \(total run ([\d.]+)', l)[0])) for l in run if 'aggregation' in l]
marks = [(m.group(1), int(m.group(2)), float(m.group(3))) for l in run for m in [re.search(r'mark (\w+)\s+frame\s+(\d+) at\s+([\d.]+)', l)] if m]
gs = int(sys.argv[2]) if len(sys.argv) > 2 else 7
print('total %.2f ms' % ag[0][2])
for c, (a, b, _) in enumerate(ag):
    fr = range(gs * c, gs * c + gs)
    g = lambda tag, fn: fn(t for (tg, f, t) in marks if tg == tag and f in fr)
    print('chunk %2d: front %7.2f-%7.2f | V %7.2f-%7.2f | back %7.2f-%7.2f | extract end %7.2f' % (
        c, g('front0', min), g('front1', max), a, b, g('back0', min), g('back1', max), g('extr1', max)))
