import csv, subprocess, sys, io
rep=sys.argv[1]; fname=sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Line No")
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
src=open(fname).read().split('\n')
# map source text -> line numbers in file (ncu line numbers are for whichever file; match by text)
agg={}
tot_s=tot_i=0
for r in rows:
    if len(r)<=iI or not r[0].isdigit(): continue
    try: ln=int(r[0]); s=int(r[iS]); i=int(r[iI])
    except ValueError: continue
    txt=r[1].strip()
    infile = ln<=len(src) and src[ln-1].strip()==txt and txt!=''
    key=(ln if infile else -1)
    agg.setdefault(key,[0,0]); agg[key][0]+=s; agg[key][1]+=i
    tot_s+=s; tot_i+=i
regions=eval(sys.argv[3])
for name,(a,b) in regions.items():
    s=sum(v[0] for k,v in agg.items() if a<=k<=b); i=sum(v[1] for k,v in agg.items() if a<=k<=b)
    print("%-28s lines %4d-%4d  samples %5.1f%%  inst %5.1f%%"%(name,a,b,100*s/tot_s,100*i/tot_i))
s,i=agg.get(-1,[0,0]); print("other files: samples %.1f%% inst %.1f%%"%(100*s/tot_s,100*i/tot_i))
