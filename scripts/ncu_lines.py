"""Per-source-line instruction counts from `ncu --page source --csv --print-source cuda,sass` for one kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
kern = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fn = fpath = None
H = None
lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": H = r; continue
    if H and fn and kern in fn and len(r) > 10 and r[0].isdigit() and r[2] == "-":  # cuda-line summary rows
        lines.append((fpath.split("/")[-1], r))
ie = H.index("Instructions Executed"); te = H.index("Thread Instructions Executed"); ss = H.index("# Samples")
tot = sum(int(r[ie] or 0) for _, r in lines); tt = sum(int(r[te] or 0) for _, r in lines); ts = sum(int(r[ss] or 0) for _, r in lines)
print("kernel filter %r: %d warp-instructions, avg threads/inst %.2f, %d samples" % (kern, tot, tt / max(tot, 1), ts))
for f, r in sorted(lines, key=lambda x: -int(x[1][ie] or 0))[:topn]:
    i = int(r[ie] or 0); t = int(r[te] or 0)
    print("%-14s %4s  inst %5.2f%%  thr/inst %5.1f  samples %5.2f%% | %s" % (f[:14], r[0], 100 * i / tot, t / max(i, 1), 100 * int(r[ss] or 0) / max(ts, 1), r[1].strip()[:100]))
