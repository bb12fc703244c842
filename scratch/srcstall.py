import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
# split into kernels
ks=[]; cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur={"name":r[1],"hdr":None,"body":[]}; ks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"]=r; continue
    if len(r)==len(cur["hdr"]): cur["body"].append(r)
which=int(sys.argv[2]) if len(sys.argv)>2 else 0
N=int(sys.argv[3]) if len(sys.argv)>3 else 45
print([k["name"][:60] for k in ks])
k=ks[which]; hdr=k["hdr"]; body=k["body"]
iS=hdr.index("# Samples"); iSrc=hdr.index("Source"); iEx=hdr.index("Instructions Executed")
stall_cols=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot=sum(int(r[iS]) for r in body)
print("total samples",tot, "instr", len(body))
top=sorted(enumerate(body), key=lambda x:-int(x[1][iS]))[:N]
for idx,r in sorted(top):
    st=sorted([(int(r[i]),hdr[i]) for i in stall_cols], reverse=True)[:2]
    print(idx, r[iS], r[iEx], r[iSrc].strip()[:80], st)
