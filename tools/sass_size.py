"""Static SASS instruction counts of the step kernel, per device function and per source function
(from `nvdisasm -gi` output).  usage: sass_size.py dis.txt [kernel-substring]"""
import bisect, collections, re, sys
dis = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else "ILi0E"
root = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
cur = sec = loc = None; grp = False
cnt, lines = collections.Counter(), collections.Counter()
for l in open(dis):
    if l.startswith("//--------------------- .text."):
        sec = l.split(".text.")[1].split()[0]; continue
    if sec is None or want not in sec: continue
    m = re.match(r"^(\$?[_A-Za-z0-9\$]+):\s*$", l.strip())
    if m and not l.strip().startswith(".L_"):
        cur = m.group(1).split("$")[-1]; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if not grp: loc = (m.group(1).split('/')[-1], int(m.group(2))); grp = True
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        cnt[cur] += 1; lines[(cur, loc)] += 1; grp = False
print("device functions:", cnt.most_common(), "total", sum(cnt.values()))
agg = collections.Counter()
for fname in ("tb_core.cuh", "tb_env.cuh", "tb_mpr.h"):
    src = open(root + "/tensegrity_rl_b200/csrc/" + fname).read().split('\n')
    fl = [(i + 1, re.search(r'(\w+)\(', l.split('TSG_FN', 1)[1].replace('_NOINLINE', '')).group(1)) for i, l in enumerate(src) if l.startswith('TSG_FN')]
    starts = [x[0] for x in fl]
    for (f, lc), n in lines.items():
        if lc and lc[0] == fname:
            k = bisect.bisect_right(starts, lc[1]) - 1
            agg[(f[-12:], fl[k][1] if k >= 0 else '?')] += n
for (f, lc), n in lines.items():
    if not lc or lc[0] not in ("tb_core.cuh", "tb_env.cuh", "tb_mpr.h"): agg[(f[-12:], str(lc[0]) if lc else None)] += n
for k, v in agg.most_common(40): print(v, k)
