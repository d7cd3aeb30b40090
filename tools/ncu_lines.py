"""Correlate an ncu SASS-page CSV (per-instruction samples / executed counts) with source lines, using the
line info nvdisasm prints for the same cubin.  usage: ncu_lines.py sass.csv dis.txt kernel_symbol [topN]"""
import csv, re, sys, collections

sass_csv, dis_txt, sym = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# --- parse nvdisasm: offset -> (func, file:line [inlined chain head])
off2loc = {}
cur_func, cur_loc, in_sec, in_group = None, None, False, False
for line in open(dis_txt):
    if line.startswith("//--------------------- .text."):
        in_sec = (".text." + sym) in line
        continue
    if not in_sec:
        continue
    m = re.match(r"^(\$?[_A-Za-z0-9\$]+):\s*$", line.strip())
    if m and not line.strip().startswith(".L_"):
        name = m.group(1)
        if "$" in name[1:]:
            name = name.split("$")[-1]
        cur_func = name
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        if not in_group:   # first line of an annotation group = innermost inlined frame
            cur_loc = (m.group(1).split("/")[-1], int(m.group(2)))
            in_group = True
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", line)
    if m:
        off2loc[int(m.group(1), 16)] = (cur_func, cur_loc, m.group(2).strip())
        in_group = False
rows = list(csv.reader(open(sass_csv)))
hdr, data = rows[1], rows[2:]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(data[0][ia], 16)
by_line, by_func = collections.Counter(), collections.Counter()
ex_line, ex_func = collections.Counter(), collections.Counter()
tot_s = tot_e = 0
for r in data:
    off = int(r[ia], 16) - base
    func, loc, ins = off2loc.get(off, ("?", None, "?"))
    s, e = int(r[isamp]), int(r[ie])
    tot_s += s; tot_e += e
    by_line[loc] += s; ex_line[loc] += e
    by_func[func] += s; ex_func[func] += e
print("total samples %d, instructions executed %d" % (tot_s, tot_e))
print("\n== by function: samples%  executed%")
for f, s in by_func.most_common():
    print("%6.2f%% %6.2f%%  %s" % (100 * s / tot_s, 100 * ex_func[f] / tot_e, f))
print("\n== top source lines: samples%  executed%  cycles/instr")
for loc, s in by_line.most_common(topn):
    e = ex_line[loc]
    print("%6.2f%% %6.2f%%  %s" % (100 * s / tot_s, 100 * e / tot_e, loc))
