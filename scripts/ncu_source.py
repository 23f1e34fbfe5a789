"""Per-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo + --import-source on):
   python scripts/ncu_source.py rep kernel-regex [top] [function-name-substring]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
only = sys.argv[4] if len(sys.argv) > 4 else None
keep = True
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg, order = {}, []
fname, idx, cur = None, None, None
names = ["# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_short_sb", "stall_barrier",
         "stall_math", "stall_wait", "stall_mio", "stall_lg", "stall_branch_resolving", "stall_no_inst", "stall_not_selected",
         "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"]
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        keep = only is None or only in r[1]
        if keep: print("== ", r[1][:120])
        cur = None
        continue
    if not keep: continue
    if r[0] == "Line No": idx = {h: k for k, h in enumerate(r) if h not in ("Source",)}; continue
    if idx is None: continue
    if r[0].isdigit():
        cur = (fname, int(r[0]), r[1].strip()[:100])
        if cur not in agg: agg[cur] = [0.0] * len(names); order.append(cur)
        continue
    if r[0] == "" and cur is not None and len(r) > 10:
        a = agg[cur]
        for k, n in enumerate(names):
            try: a[k] += float(r[idx[n]])
            except Exception: pass
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print("total samples %d, warp-instructions %d, avg threads/inst %.1f" % (tot, toti, sum(a[2] for a in agg.values()) / toti))
for key in sorted(agg, key=lambda k: -agg[k][0])[:top]:
    a = agg[key]
    print("%5.1f%% smp %5.1f%% ins thr %4.1f | %s:%d %s\n        [long %d short %d bar %d math %d wait %d mio %d lg %d br %d noinst %d notsel %d | smem wf %d ideal %d]" %
          (100 * a[0] / tot, 100 * a[1] / toti, a[2] / max(a[1], 1), key[0], key[1], key[2], *a[3:]))
