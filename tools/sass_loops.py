"""list the inner loops of a SASS dump that contain FP64 reciprocals, with their instruction mix"""
import re, sys, collections
lines=[l for l in open(sys.argv[1]) if re.search(r'/\*[0-9a-f]{4,}\*/',l)]
ins=[]
for l in lines:
    m=re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
addr={a:i for i,(a,_) in enumerate(ins)}
for i,(a,t) in enumerate(ins):
    if 'BRA' in t:
        m2=re.search(r'0x([0-9a-f]+)',t)
        if m2:
            tgt=int(m2.group(1),16)
            if tgt<a and tgt in addr:
                body=ins[addr[tgt]:i+1]
                nm=sum('MUFU.RCP64H' in x for _,x in body)
                if nm>=int(sys.argv[2] if len(sys.argv)>2 else 4) and len(body)<int(sys.argv[3] if len(sys.argv)>3 else 600):
                    c=collections.Counter((x.split()[1] if x.startswith('@') else x.split()[0]).split('.')[0] for _,x in body)
                    dp=sum(v for k,v in c.items() if k in('DFMA','DADD','DMUL','DSETP'))
                    print("loop",hex(tgt),hex(a),"len",len(body),"mufu",nm,"dp",dp,dict(c.most_common(16)), 'LDL', c.get('LDL',0), 'STL', c.get('STL',0))
