"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K` per CUDA source line:
share of warp-stall samples, share of executed warp instructions, dominant stall reasons."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fname = ""
    hdr = None
    out = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        if r[2] != "-":  # SASS row
            continue
        try:
            smp = int(r[hdr.index("# Samples")])
            ins = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h and i < len(r):
                try:
                    v = int(r[i])
                except ValueError:
                    v = 0
                if v:
                    stalls[h[6:]] = v
        out.append((smp, ins, fname, r[0], r[1].strip(), stalls))
    ts = sum(o[0] for o in out) or 1
    ti = sum(o[1] for o in out) or 1
    print(f"total samples {ts}, warp instructions {ti}")
    for smp, ins, f, ln, src, st in sorted(out, key=lambda o: -o[0])[:top]:
        s3 = ",".join(f"{k}:{v * 100 // max(smp, 1)}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{smp * 100 / ts:5.1f}%smp {ins * 100 / ti:5.1f}%ins {f}:{ln:>4} [{s3}] {src[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
