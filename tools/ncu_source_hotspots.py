import re,csv,collections,subprocess,sys
sass_file, rep, kregex = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv)>4 else 40
lines=open(sass_file).read().split('\n')
cur=None
insts=[]
for l in lines:
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
    if m:
        insts.append((int(m.group(1),16),m.group(2),cur))
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+kregex],capture_output=True,text=True).stdout
rows=list(csv.reader(out.split('\n')))
segs=[]; 
for r in rows:
    if r and r[0]=='Kernel Name': segs.append({'name':r[1],'hdr':None,'data':[]}); continue
    if not segs: continue
    if segs[-1]['hdr'] is None: segs[-1]['hdr']=r; continue
    if len(r)>5: segs[-1]['data'].append(r)
seg=segs[0]
hdr=seg['hdr']; data=seg['data']
print(seg['name'][:60], len(data), len(insts))
ia=hdr.index('Address'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base=int(data[0][ia],16)
byline=collections.Counter(); exline=collections.Counter(); tot=0; stl=collections.defaultdict(collections.Counter)
for k,r in enumerate(data):
    s=int(r[isamp]); tot+=s
    cur=insts[k][2] if k<len(insts) else None
    byline[cur]+=s; exline[cur]+=int(r[iex])
    for i in stall_cols:
        v=int(r[i]) if r[i] else 0
        if v: stl[cur][hdr[i]]+=v
print('total samples',tot)
for cur,s in byline.most_common(top):
    f,l=cur if cur else ('?',0)
    top3=', '.join(f"{k[6:]}={v}" for k,v in stl[cur].most_common(3))
    print(f"{f}:{l:4d} {s:7d} {100*s/tot:5.1f}%  inst={exline[cur]:9d}  {top3}")
