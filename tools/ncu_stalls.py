"""Per source line of one file: stall samples with the dominant stall reasons (from an ncu report).
usage: python tools/ncu_stalls.py report.ncu-rep file.cuh [lo hi] [top]"""
import csv, subprocess, sys

def main():
    rep, fname = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    cur, hdr, lines, total = "", None, [], 0
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r; continue
        if hdr and len(r) == len(hdr) and r[0] != "" and r[2] == "-":
            d = dict(zip(hdr, r))
            smp = int(d["# Samples"] or 0)
            total += smp
            if cur == fname and lo <= int(r[0]) <= hi:
                reasons = sorted(((int(d[k] or 0), k[6:]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:3]
                lines.append((smp, int(r[0]), r[1].strip()[:70], int(d["Instructions Executed"] or 0), reasons))
    sub = sum(l[0] for l in lines)
    print("samples in range: %d of %d (%.1f%%)" % (sub, total, 100.0 * sub / max(total, 1)))
    for smp, ln, src, inst, reasons in sorted(lines, reverse=True)[:top]:
        print("%5.1f%% %5d  L%-4d %-70s  %s" % (100.0 * smp / max(total, 1), inst, ln, src, " ".join("%s:%d" % (k, v) for v, k in reasons if v)))
main()
