import csv, collections, sys
# usage: ncu_region.py file.csv srcfile lo hi  -> stall reason totals for source lines [lo,hi]
rows=list(csv.reader(open(sys.argv[1])))
fn, lo, hi = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
hdr=None; cur_file=None; cur_line=None
tot=collections.Counter(); ninst=0; nexec=0
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if len(r)<8: continue
    if r[0] not in ('',) and r[2]=='-':
        cur_line=(cur_file,int(r[0])); continue
    if r[0]=='' and r[2].startswith('0x') and cur_line and cur_line[0]==fn and lo<=cur_line[1]<=hi:
        ninst+=1
        try: nexec+=int(r[7])
        except: pass
        for i,h in enumerate(hdr):
            if h.startswith('stall_') and 'Not Issued' not in h:
                try: tot[h]+=int(r[i])
                except: pass
print("region %s:%d-%d  sass instrs %d  executed %d  samples %d"%(fn,lo,hi,ninst,nexec,sum(tot.values())))
for k,v in tot.most_common(8): print("   %-24s %d"%(k,v))
