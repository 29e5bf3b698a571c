"""Per-SASS-instruction memory traffic of an .ncu-rep captured with --import-source on: shared wavefronts, global tag
requests and L2 sectors, with the stall reasons sampled at the instruction.
usage: python profiles/sass_mem_hot.py file.ncu-rep [topN] [kernel-regex]"""
import csv
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, agg = None, []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            if agg:
                break                      # first kernel only
            continue
        if not hdr or len(r) != len(hdr):
            continue
        g = lambda n: float(r[hdr.index(n)] or 0)
        agg.append((g("L1 Wavefronts Shared"), g("L1 Tag Requests Global"), g("L2 Theoretical Sectors Global"),
                    g("Instructions Executed"), g("# Samples"), g("stall_long_sb"), g("stall_lg"), g("stall_mio"),
                    g("stall_short_sb"), g("stall_wait"), g("stall_math"), g("stall_not_selected"), r[1].strip()[:70]))
    tot = [sum(a[i] for a in agg) for i in range(12)]
    print("totals: shared wavefronts %.3g, global tag requests %.3g, L2 sectors %.3g, instructions %.3g, samples %d"
          % (tot[0], tot[1], tot[2], tot[3], tot[4]))
    print("stall samples: long_sb %d lg %d mio %d short_sb %d wait %d math %d not_selected %d" % tuple(tot[5:12]))
    print("%10s %10s %10s %10s %7s  %s" % ("shr_wave", "glb_tag", "l2_sect", "inst", "samples", "sass"))
    key = lambda a: a[0] + a[1] + a[4] * tot[0] / max(tot[4], 1) * 0.0
    for a in sorted(agg, key=lambda a: -(a[0] + a[1]))[:top]:
        print("%10.3g %10.3g %10.3g %10.3g %7d  %s" % (a[0], a[1], a[2], a[3], a[4], a[12]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
