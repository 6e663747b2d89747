import sys, collections
rows=[l.split() for l in open(sys.argv[1]) if l.startswith("FT ")]
rows=[[int(x) for x in r[1:]] for r in rows]
# group launches: consecutive blocks of 148 (by order of appearance is unreliable) -> cluster by time
rows.sort(key=lambda r:r[1])
G=148
launch=rows[-G:]
t0=min(r[1] for r in launch)
names=["enter","bar1done","reduced","arrive2","bar2done","normed","adamdone"]
for i,n in enumerate(names):
    v=[r[1+i]-t0 for r in launch]
    print("%-9s min %6d  median %6d  max %6d ns   (argmax cta %d)"%(n,min(v),sorted(v)[len(v)//2],max(v),launch[v.index(max(v))][0]))
