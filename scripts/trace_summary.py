"""Summarise a JA_SC_TRACE=1 log of scripts/pass_times.py: per sumcheck shape, where the host time of the LAST pass went."""
import collections, re, sys
lines = open(sys.argv[1]).read().splitlines()
idx = [i for i, l in enumerate(lines) if l.startswith('pass') or l.startswith("{'pass'")]
seg = lines[idx[-2] + 1:idx[-1]]
pat = re.compile(r'\[sc n=(\d+) rounds=(\d+)\] cumulative us: launch=(\d+) inv=(\d+) wait\+interp=(\d+) transcript=(\d+) ingest=(\d+)')
pc = re.compile(r'\[sc-call n=(\d+)\] build=(\d+) us loop=(\d+) us release=(\d+) us')
prev = None
for l in lines[:idx[-2]][::-1]:
    m = pat.match(l)
    if m:
        prev = list(map(int, m.groups()[2:])); break
agg = collections.defaultdict(lambda: [0] * 10)
key = None
for l in seg:
    m = pat.match(l)
    if m:
        n, r = int(m.group(1)), int(m.group(2)); cur = list(map(int, m.groups()[2:]))
        d = [a - b for a, b in zip(cur, prev)]; prev = cur
        key = (n, r); a = agg[key]; a[0] += 1; a[1] += r
        for i in range(5): a[2 + i] += d[i]
        continue
    m = pc.match(l)
    if m and key is not None:
        a = agg[key]
        a[7] += int(m.group(2)); a[8] += int(m.group(3)); a[9] += int(m.group(4))
print("(n,rounds) calls | launch inv wait transcript ingest | build loop release (us, summed) | loop us/round")
tot = [0] * 8
for k, a in sorted(agg.items()):
    print(k, a[0], a[2:7], a[7:10], round(a[8] / max(a[1], 1), 1))
    for i in range(8): tot[i] += a[2 + i]
print("totals", tot)
print(lines[idx[-1]])
