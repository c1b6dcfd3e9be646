"""Per source line: warp instructions executed and stall samples of the first kernel in an `ncu -i X.ncu-rep --page source --csv
--print-source cuda,sass` export (argv[1]); tolerant of source text with quotes and commas (columns are read from the right).
argv[2] = number of lines to list."""
import sys, collections, re
agg=collections.OrderedDict(); cur=None
NC=66; S_OFF=60; I_OFF=59          # defaults of the first export this was written for; replaced by each table's header
for line in open(sys.argv[1]):
    line=line.rstrip('\n')
    if line.startswith('"File Path"'):
        cur=line.split('","')[1].strip('"').split('/')[-1]; continue
    if not line.startswith('"') : continue
    parts=line.split('","')
    if parts[0].strip('"') == 'Line No':          # header of a function's table: column positions, counted from the right
        hdr=[h.strip('"') for h in parts]
        NC=len(hdr); S_OFF=NC-hdr.index('# Samples'); I_OFF=NC-hdr.index('Instructions Executed')
        continue
    if len(parts)<NC: continue
    first=parts[0].strip('"')
    if not first.isdigit(): continue
    try:
        samples=int(parts[-S_OFF]); inst=int(parts[-I_OFF])
    except: continue
    src='","'.join(parts[1:len(parts)-NC+2])[:110]
    key=(cur,int(first))
    a=agg.get(key,(0,0,''))
    agg[key]=(a[0]+inst,a[1]+samples,src)
tot=sum(v[0] for v in agg.values()); tots=sum(v[1] for v in agg.values())
print('inst',tot,'samples',tots)
by=collections.Counter(); bys=collections.Counter()
for (f,l),v in agg.items(): by[f]+=v[0]; bys[f]+=v[1]
for f,v in by.most_common(): print('%-20s inst %5.1f%% samples %5.1f%%'%(f,100*v/tot,100*bys[f]/tots))
top=int(sys.argv[2]) if len(sys.argv)>2 else 40
print('--- top lines by instructions')
for (f,l),v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print('%-14s %5d inst %5.2f%% smp %5.2f%%  %s'%(f,l,100*v[0]/tot,100*v[1]/tots,v[2]))
print('--- top lines by samples')
for (f,l),v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:top]:
    print('%-14s %5d inst %5.2f%% smp %5.2f%%  %s'%(f,l,100*v[0]/tot,100*v[1]/tots,v[2]))
