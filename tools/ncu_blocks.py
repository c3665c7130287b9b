"""Basic-block view of an ncu report's SASS page: executed warp instructions, share, average active lanes, stall samples.
Usage: python tools/ncu_blocks.py report.ncu-rep [min_share_percent] [launch index in the report, default 0]"""
import csv, subprocess, sys
path = sys.argv[1]; min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
start = starts[which]; end = starts[which + 1] - 1 if which + 1 < len(starts) else len(rows)
hdr = rows[start]
if start > 0 and rows[start - 1] and rows[start - 1][0] == 'Kernel Name': print('kernel', rows[start - 1][1])
ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed'); ss = hdr.index('# Samples')
blocks = []; cur = None; tot = 0; tot_t = 0
for r in rows[start + 1:end]:
    try: i_ = int(r[ie]); t_ = int(r[te]); s_ = int(r[ss])
    except Exception: continue
    tot += i_; tot_t += t_
    op = r[1].strip()
    last = cur['ops'][-1] if cur else ''
    ends = any(k in last for k in ('BRA', 'BSYNC', 'EXIT', 'RET', 'CALL', 'WARPSYNC'))
    if cur and cur['cnt'] == i_ and not ends:
        cur['n'] += 1; cur['inst'] += i_; cur['thr'] += t_; cur['s'] += s_; cur['ops'].append(op)
    else:
        cur = {'addr': r[0], 'cnt': i_, 'n': 1, 'inst': i_, 'thr': t_, 's': s_, 'ops': [op]}; blocks.append(cur)
print('total warp inst', tot, 'thread inst', tot_t, 'avg lanes %.2f' % (tot_t / max(tot, 1)))
for b in blocks:
    if b['inst'] < tot * min_share / 100: continue
    ops = ' '.join((o.split()[1] if o.startswith('@') else o.split()[0]) for o in b['ops'])
    print(f"{b['addr'][-5:]} n={b['n']:3d} exec={b['cnt']/1e3:8.0f}k inst={100*b['inst']/tot:5.1f}% thr={100*b['thr']/tot_t:5.1f}% avg={b['thr']/max(b['inst'],1):5.1f} smp={b['s']:5d} | {ops[:140]}")
