"""Aggregate ncu warp-stall samples per CUDA source line.
   python tools/ncu_lines.py report.ncu-rep [function-substr] [top-n]"""
import csv, subprocess, sys
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = cur_fn = None; hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or sub not in (cur_fn or ""): continue
    if r[0].isdigit():
        si = hdr.index("# Samples")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        n = int(r[si]) if r[si].isdigit() else 0
        key = (cur_fn[:60], cur_file.split("/")[-1], int(r[0]))
        st = {hdr[i][6:]: (int(r[i]) if r[i].isdigit() else 0) for i in stall_cols}
        if key in agg:
            agg[key][0] += n
            for k, v in st.items(): agg[key][2][k] = agg[key][2].get(k, 0) + v
        else:
            agg[key] = [n, r[1], st]
fns = sorted(set(k[0] for k in agg))
for fn in fns:
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    tot = sum(v[0] for _, v in items)
    print("==", fn, "total samples", tot)
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:topn]:
        st = sorted(v[2].items(), key=lambda kv: -kv[1])[:2]
        print(f"{v[0]:6d} {k[1]}:{k[2]:<4d} {v[1].strip()[:88]:<88} {st}")
