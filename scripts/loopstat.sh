# usage: scripts/loopstat.sh [lib.so] [nodes-per-lane 1|2]: instruction mix of the mixture kernel's hot row loop
pat="pw_frontier_kernel2INS_12MixtureModelILi8ELi2EdEEdLi${2:-1}E"
cuobjdump -sass ${1:-fetching_b200/lib/libpfetch.so} | awk -v pat="$pat" '$0 ~ "Function : .*" pat {f=1} f{print}' | awk 'NR>1 && /Function :/{exit} {print}' > /tmp/mix.sass
python3 - <<'PY'
import re
lines=[l for l in open('/tmp/mix.sass') if re.search(r'/\*[0-9a-f]{4}\*/',l)]
ins=[]
for l in lines:
    m=re.search(r'/\*([0-9a-f]{4})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16), m.group(2).strip()))
addr={a:i for i,(a,_) in enumerate(ins)}
# find backward branches; choose loop with most DFMA
best=None
for i,(a,t) in enumerate(ins):
    m=re.search(r'BRA\s+(?:[A-Z0-9!]+,\s*)?0x([0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a and tgt in addr:
            body=ins[addr[tgt]:i+1]
            nd=sum(1 for _,x in body if re.search(r'\bD(FMA|ADD|MUL)\b',x))
            ok=any('I2F' in x for _,x in body) and not any('CALL' in x for _,x in body) and not any('BAR' in x for _,x in body)
            if ok and nd>=60 and (best is None or len(body)<len(best[1])): best=(nd,body)
nd,body=best
from collections import Counter
c=Counter()
for _,x in body:
    x=re.sub(r'^@!?U?P\d\s+','',x)
    c[x.split()[0].split('.')[0]]+=1
f=sum(v for k,v in c.items() if k in('DFMA','DADD','DMUL'))
print("loop instrs",len(body),"FP64",f,"other",len(body)-f,"model cycles",2*f+len(body)-f)
print(dict(c.most_common()))
PY
