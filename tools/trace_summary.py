import re,sys
txt=open(sys.argv[1]).read().split('TRACE ')
for blk in txt[1:]:
    lines=blk.split('\n'); hdr=lines[0]
    for l in lines[2:9]:
        name=l.split()[0] if l.split() else ''
        vals=[int(v) for v in l.split()[1:] if re.fullmatch(r'-?\d+',v)]
        if name=='mma_got_full':
            d=sorted(vals[i+1]-vals[i] for i in range(len(vals)-1))
            med=d[len(d)//2] if d else 0
        if name=='entry_setup_proddone_alldone': tot=vals[-1]
    print(hdr[:70], '| median K-block cadence', med, '| CTA0 total clk', tot)
