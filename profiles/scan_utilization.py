import csv,sys
from collections import defaultdict
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and not r[0].startswith("==")]
hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
d=defaultdict(dict)
for r in rows[1:]: d[(r[iid], r[ik][:40])][r[im]]=float(r[iv].replace(",",""))
seen=set()
for (i,k),v in d.items():
    if k in seen: continue
    seen.add(k)
    print("%-42s %9.1f us  act/elapsed %.2f  issue %3.0f%%  dram %3.0f%%  l1 %3.0f%%"%(k, v["gpu__time_duration.sum"]/1e3, v["smsp__cycles_active.avg"]/v["sm__cycles_elapsed.max"], v["smsp__issue_active.avg.pct_of_peak_sustained_active"], v["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"], v["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]))
