import sys
f=sys.argv[1]
L=open(f).read().splitlines()
print(L[0])
d={}
for l in L[1:]:
    a=l.split(); r,it=int(a[0]),int(a[1]); d[(r,it)]=[int(x) for x in a[2:]]
t0=min(x for v in d.values() for x in v if x>0)
def g(r,it,e):
    x=d[(r,it)][e]; return (x-t0) if x>0 else None
n=max(it for (r,it),v in d.items() if any(v))+1
print("items",n)
names={0:["p0wait","p0go","p1wait","p1go"],1:["start","a1e","s0","c0","s1","c1","aE","st"],2:["m1w","m1go","m1is","m2w","m2go","m2is"],3:["e1w","e1go","e1dn","e2s","e2go","e2dn"]}
for it in range(min(n,int(sys.argv[2]) if len(sys.argv)>2 else 10)):
    for r in range(4):
        s=" ".join(f"{names[r][e]}={g(r,it,e)}" for e in range(len(names[r])) if g(r,it,e) is not None)
        print(f"it{it} r{r}: {s}")
# period
e2=[g(3,it,5) for it in range(n)]
print("e2 done deltas", [e2[i+1]-e2[i] for i in range(min(n-1,21))])
