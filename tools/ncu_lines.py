"""Summarise an ncu report per CUDA source line: instructions executed and stall samples.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file = ""
    hdr = None
    lines = []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr and len(r) > 5 and r[0] != "" and r[2] == "-":
            def g(name):
                try:
                    return int(r[hdr[name]] or 0)
                except (KeyError, ValueError):
                    return 0
            lines.append((cur_file, int(r[0]), r[1].strip()[:90], g("Instructions Executed"), g("# Samples"),
                          g("Warp Stall Sampling (Not-issued Samples)")))
    tot_i = sum(l[3] for l in lines)
    tot_s = sum(l[4] for l in lines)
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    print("--- by instructions")
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * l[3] / tot_i, 100 * l[4] / max(tot_s, 1), l[0], l[1], l[2]))
    print("--- by stall samples")
    for l in sorted(lines, key=lambda l: -l[4])[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * l[4] / max(tot_s, 1), 100 * l[3] / tot_i, l[0], l[1], l[2]))


if __name__ == "__main__":
    main()
