import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 40
cur_file=None; cur_line=None; src={}
agg=collections.defaultdict(lambda:[0,0])
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if len(r)<8: continue
    if r[0] not in ('','Line No') and r[2]=='-':
        cur_line=(cur_file,int(r[0])); src[cur_line]=r[1]; continue
    if r[0]=='' and r[2].startswith('0x'):
        try:
            agg[cur_line][0]+=int(r[4]); agg[cur_line][1]+=int(r[7])
        except: pass
tot=sum(v[0] for v in agg.values()); toti=sum(v[1] for v in agg.values())
print("total samples",tot,"total inst",toti)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s"%(100*v[0]/tot,100*v[1]/toti,k[0],k[1],src[k][:100]))
