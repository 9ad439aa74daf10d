import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); im=h.index("Metric Name"); iv=h.index("Metric Value"); iid=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[iid]), r[ik]),{})[r[im]]=float(r[iv].replace(",",""))
n=float(sys.argv[2]); step=int(sys.argv[3])
for (i,k),m in sorted(d.items()):
    if i%step==step-1:
        print(i, k[:16], "ms=%.2f"%(m["gpu__time_duration.sum"]/1e6), "dramR B/op=%.1f"%(m["dram__bytes_read.sum"]/n), "dramW B/op=%.1f"%(m["dram__bytes_write.sum"]/n))
