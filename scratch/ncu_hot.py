"""hot-code footprint per source line from `ncu --page source --csv --print-source cuda,sass`:
python scratch/ncu_hot.py file.csv <executions-threshold> [top]"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
thr=float(sys.argv[2]); top=int(sys.argv[3]) if len(sys.argv)>3 else 40
cur_file=None; cur_line=None; src={}
hot=collections.Counter(); ex=collections.Counter(); stat=collections.Counter(); samp=collections.Counter()
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if len(r)<8: continue
    if r[0] not in ('','Line No') and r[2]=='-':
        cur_line=(cur_file,int(r[0])); src[cur_line]=r[1]; continue
    if r[0]=='' and r[2].startswith('0x'):
        try: e=int(r[7]); sa=int(r[4])
        except: continue
        stat[cur_line]+=1; samp[cur_line]+=sa
        if e > thr: hot[cur_line]+=1; ex[cur_line]+=e
tote=sum(ex.values()); tots=sum(samp.values())
print("static total", sum(stat.values()), "hot static", sum(hot.values()), "=> %.0f KB"%(sum(hot.values())*16/1024))
for k,v in hot.most_common(top):
    print("%4d hot / %4d static  exec %5.1f%%  samp %5.1f%%  %s:%d  %s"%(v,stat[k],100*ex[k]/tote,100*samp[k]/tots,k[0],k[1],src[k][:80]))
