"""Per-source-line totals of an .ncu-rep captured with --import-source on: warp instructions, lanes, stall samples.
usage: python profiles/source_hot.py file.ncu-rep [topN]"""
import csv
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, agg, total = "", None, [], 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            # the source text holds unescaped quotes: count the numeric columns from the right
            def col(name):
                return int(r[hdr.index(name) - len(hdr)])
            ie, te, sm = col("Instructions Executed"), col("Thread Instructions Executed"), col("# Samples")
            agg.append((ie, te, sm, cur_file, int(r[0]), r[1].strip()[:90]))
            total += ie
    tot_s = sum(a[2] for a in agg)
    print("total warp instructions %d, samples %d" % (total, tot_s))
    for ie, te, sm, f, ln, src in sorted(agg, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% smp %4.1f lanes  %s:%d  %s" % (100.0 * ie / total, 100.0 * sm / max(tot_s, 1),
                                                              te / max(ie, 1), f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
