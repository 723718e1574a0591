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
# quad contraction kernel (only in -DSC_TRACE_QUAD builds): absolute clocks relative to the kernel's first instruction
for s in [s for s in sections if 'quad' in s[0]][-3:]:
    sec = s[1]
    t0 = sec.get((0, 0), (0, 0, 0))[0]
    if t0 == 0: continue
    print('==', s[0], ' role 0 = loader warp (i=0: start/prologue done/pdl done; i=1+rd: A issued/A scattered/B scattered),'
          ' role 1 = MMA+epilogue (i=1+rd: tiles full/MMAs issued; i=5: D full/stores issued/all warps done)')
    for r in range(2):
        row = []
        for i in range(0, 6):
            a, b, c = sec.get((r, i), (0, 0, 0))
            if a == 0: continue
            row.append(f"{i}:[{a-t0:6d} {b-t0 if b else 0:6d} {c-t0 if c else 0:6d}]")
        print(r, ' '.join(row))
