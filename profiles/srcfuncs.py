"""Per CUDA source line and kernel: share of executed warp instructions, of stall samples, active threads per warp
instruction. Input: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > file.csv`.
    python profiles/srcfuncs.py file.csv [top] [kernel substring]"""
import csv
import sys


def main(path, top=25, only=""):
    rows = list(csv.reader(open(path)))
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
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h and i < len(r):
                try:
                    v = int(r[i])
                except ValueError:
                    v = 0
                if v:
                    st[h[6:]] = v
        funcs[cur].append((ins, smp, thr, fname, r[0], r[1].strip(), st))
    for f, out in funcs.items():
        if only and only not in f:
            continue
        ti = sum(o[0] for o in out) or 1
        ts = sum(o[1] for o in out) or 1
        print("=====", f[:90], "warp instr", ti, "samples", ts)
        for ins, smp, thr, fn, ln, src, st in sorted(out, key=lambda o: -o[1])[:top]:
            s3 = ",".join(f"{k}:{v * 100 // max(smp, 1)}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print(f"{smp * 100 / ts:5.1f}%smp {ins * 100 / ti:5.1f}%ins thr/warp {thr / max(ins, 1):4.1f} {fn}:{ln:>4} [{s3}] {src[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25, sys.argv[3] if len(sys.argv) > 3 else "")
