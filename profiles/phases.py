"""Instruction share per phase of the tile-sweep kernels (line ranges of pfd_tilesweep.cuh), from the ncu source page.
    python profiles/phases.py file.csv <kernel substring>"""
import bisect
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
funcs, cur, fname, hdr = {}, None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur = r[1]
        funcs.setdefault(cur, [])
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or cur is None or len(r) < 10 or r[2] != "-":
        continue
    try:
        smp = int(r[hdr.index("# Samples")])
        ins = int(r[hdr.index("Instructions Executed")])
        thr = int(r[hdr.index("Thread Instructions Executed")])
    except ValueError:
        continue
    funcs[cur].append((ins, smp, thr, fname, int(r[0]) if r[0].isdigit() else -1))
src = open("/root/repo/pyflwdir_b200/csrc/pfd_tilesweep.cuh").read().splitlines()


TEXT = "\n".join(src)


def find(pat):
    k = TEXT.find(pat)
    return TEXT.count("\n", 0, k) + 1 if k >= 0 else None


names = [("stage_graph", "void ts_stage_graph"), ("load16", "void ts_load16"), ("store16", "void ts_store16"),
         ("activate", "void ts_activate"), ("act_bit", "uint32_t ts_act_bit"), ("scan_word", "void ts_scan_word"),
         ("live/halo/popc", "ts_live4(uint32_t"), ("push", "void ts_push"), ("structs", "struct TsShared"), ("finish", "void ts_finish"),
         ("up_step", "int ts_up_step"), ("up head+stage", "void ts_up_visit"), ("up records", "// per-cell records (4 cells per word of the planes)"),
         ("up fill q", "// In-tile dataflow, level by level"), ("up rounds", "    uint32_t lo = 0;\n    for (int rd = 0;; ++rd) {\n        ts_sync<NT>();\n        const uint32_t n"),
         ("up kernel loop", "// All passes in one cooperative launch"), ("ops", "// streams.accuflux (up): accu"),
         ("down head", "void ts_down_visit"), ("down records+roots", "// records: unresolved children inside the tile"),
         ("down rounds", "// rounds, level by level: a lane takes ONE resolved cell"), ("down kernel", "tile_down_sweep_kernel(TsArgs"),
         ("hand op", "struct HandTileOp")]
marks = sorted([(n, find(p)) for n, p in names if find(p)], key=lambda m: m[1])
for f, out in funcs.items():
    if want not in f:
        continue
    ti = sum(o[0] for o in out)
    print(f[:70], "total warp instr", ti)
    agg = {}
    for ins, smp, thr, fn, ln in out:
        if fn != "pfd_tilesweep.cuh":
            key = "other:" + fn
        else:
            i = bisect.bisect_right([m[1] for m in marks], ln) - 1
            key = marks[i][0] if i >= 0 else "top"
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += ins
        a[1] += smp
        a[2] += thr
    for k, (ins, smp, thr) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
        print(f"  {k:30s} {ins * 100 / ti:5.1f}% ins  thr/warp {thr / max(ins, 1):5.1f}  samples {smp}")
