"""Summarise an event trace of the decoder stream kernel (tests/cuda/dec_stream_trace traceN > trace.txt):
per layer the MMA warp's span and waits, and optionally the per-run timeline of selected layers.
usage: python tools/trace_report.py trace.txt [layer ...]"""
import collections
import sys

ev = collections.defaultdict(list)
for line in open(sys.argv[1]):
    if not line.startswith('TRACE'):
        continue
    _, role, i, a, b, t1, t2, t3 = line.split()
    ev[role].append((int(i), int(a), int(b), int(t1), int(t2), int(t3)))
t0 = min(e[3] for e in ev['mma'] if e[3])
R = lambda t: round((t - t0) / 1000.0, 1)
detail = [int(x) for x in sys.argv[2:]]
bylayer = collections.defaultdict(list)
for e in ev['mma']:
    bylayer[e[1]].append(e)
prev_end = None
tot_span = tot_acc = tot_af = tot_gap = 0
print("layer runs span_kcyc acc_wait afull_wait gap_from_prev")
for l in sorted(bylayer):
    es = sorted(bylayer[l])
    start, end = es[0][3], es[-1][5]
    accw = sum(e[4] - e[3] for e in es)
    afw = sum(e[5] - e[4] for e in es)
    gap = (start - prev_end) / 1000 if prev_end else 0
    tot_span += (end - start) / 1000; tot_acc += accw / 1000; tot_af += afw / 1000; tot_gap += gap
    if not detail:
        print(l, len(es), round((end - start) / 1000, 1), round(accw / 1000, 1), round(afw / 1000, 1), round(gap, 1))
    prev_end = end
print("TOTAL kcyc: span %.0f (acc_wait %.0f, afull_wait %.0f) + gaps %.0f" % (tot_span, tot_acc, tot_af, tot_gap))
for L in detail:
    print('--- layer', L)
    pr = sorted(e for e in ev['producer'] if e[1] == L)
    mm = sorted(e for e in ev['mma'] if e[1] == L)
    for p, m in zip(pr, mm):
        print('run kk=%3d | prod: start %7.1f flags_ok +%5.1f slot_ok +%5.1f | mma: arrive %7.1f accwait %4.1f afull_wait %5.1f got %7.1f' % (
            p[2], R(p[3]), (p[4] - p[3]) / 1000, (p[5] - p[4]) / 1000, R(m[3]), (m[4] - m[3]) / 1000, (m[5] - m[4]) / 1000, R(m[5])))
