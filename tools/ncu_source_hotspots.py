import csv,sys
from collections import Counter
allrows=list(csv.reader(open(sys.argv[1])))
# split into sections
secs=[];cur=None
for r in allrows:
    if r and r[0]=="Kernel Name": cur={'name':r[1],'rows':[]}; secs.append(cur); continue
    if cur is not None: cur['rows'].append(r)
want=sys.argv[3] if len(sys.argv)>3 else None
for sec in secs:
    if want and want not in sec['name']: continue
    rows=[r for r in sec['rows'] if len(r)>10]
    hdr=rows[0]; body=[r for r in rows[1:] if r[hdr.index('# Samples')].isdigit()]
    ia=hdr.index('Source'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
    tot=sum(int(r[isamp]) for r in body)
    print("=====",sec['name'],"total samples",tot,"n sass",len(body), "instr executed", sum(int(r[iex]) for r in body))
    top=sorted(body, key=lambda r:-int(r[isamp]))[:int(sys.argv[2]) if len(sys.argv)>2 else 30]
    for r in top: print(r[isamp], r[iex], r[ia][:100])
    c=Counter(); cs=Counter()
    for r in body:
        t=r[ia].split()
        op=t[1] if t[0].startswith('@') else t[0]
        op=op.split('.')[0]
        c[op]+=int(r[iex]); cs[op]+=int(r[isamp])
    print(c.most_common(25))
    print(cs.most_common(12))
    break
