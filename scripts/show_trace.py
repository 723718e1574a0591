import sys
sections=[]; cur=None
for line in open(sys.argv[1]):
    if line.startswith('#'):
        cur=[line.strip(), {}]; sections.append(cur); continue
    r,i,a,b,c=map(int,line.split()); cur[1][(r,i)]=(a,b,c)
for name in ('analysis','synthesis'):
    secs=[s for s in sections if name in s[0]]
    if not secs: continue
    sec=secs[-1][1]
    t0=min(v[0] for v in sec.values() if v[0]>0)
    print('==',name,'(role: 0 loader/prep 1 S1/SA 2 epi1/epiA 3 S2/SB 4 epi2/epiB)  i:[start wWAIT xWORK]')
    for r in range(8):
        row=[]
        for i in range(2,10):
            a,b,c=sec.get((r,i),(0,0,0))
            if a==0: continue
            row.append(f"{i}:[{a-t0:6d} w{b-a:5d} x{c-b:5d}]")
        print(r,' '.join(row))
