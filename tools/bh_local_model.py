"""CPU model of the combinatorics of bh_emit_local_kernel + the climb (nbodysim_b200/csrc/barnes_hut.cu), on SYMBOLIC values:
which cells a CTA window finishes locally, which children it adds in which order, which (child, parent) edges it hands to the
top of the tree, what the skip pointers are -- for arbitrary window / halo sizes, duplicates and deep chains -- checked against a
plain recursion over the sorted keys (tests/test_bh_build_models.py).  Also the 8-ary owner search (bh_first_with_prefix).
The arithmetic itself is covered on the GPU against the oracle; this model exists so that the index logic of the kernel can be
changed and re-checked in a second on a machine without a GPU."""
import random, sys
LEVELS=32; BITS=2
def lcp(a,b):
    if a==b: return LEVELS
    x=a^b
    return (64 - x.bit_length())//BITS
def build_ref(keys):
    """reference: cells as dict (owner s, depth d) -> tuple of child cell ids in order / ('leaf', s)"""
    n=len(keys)
    A=[-1]+[lcp(keys[i-1],keys[i]) for i in range(1,n)]+[-1]
    first=[0]*n; leaf=[0]*n; count=[0]*n
    for s in range(n):
        if s>0 and keys[s-1]==keys[s]: continue
        lp=A[s]; t=s+1
        while t<n and keys[t]==keys[s]: t+=1
        ln = lcp(keys[s],keys[t]) if t<n else -1
        leafd = 0 if (lp<0 and ln<0) else max(lp,ln)+1
        first[s]=lp+1; leaf[s]=leafd; count[s]=leafd-first[s]+1
    offs=[0]*(n+1)
    for s in range(n): offs[s+1]=offs[s]+count[s]
    return A,first,leaf,count,offs
def ref_values(keys,A,first,leaf,count,offs):
    n=len(keys)
    val={}; nxt={}
    # process cells bottom up by recursion over spans
    def end_of(s,d):
        t=s+1
        while t<n and A[t]>=d: t+=1
        return t
    import functools
    sys.setrecursionlimit(10000)
    def cellval(s,d):
        c=offs[s]+d-first[s]
        e=end_of(s,d)
        nxt[c]= offs[e] if e<n else 0
        if d==leaf[s]:
            v=('leaf',s); val[c]=v; return v
        ch=[cellval(s,d+1)]
        t=end_of(s,d+1)
        while t<e:
            assert A[t]==d, (s,d,t,A[t])
            ch.append(cellval(t,d+1)); t=end_of(t,d+1)
        v=tuple(ch); val[c]=v; return v
    cellval(0,0)
    return val,nxt
def model(keys,M,H):
    n=len(keys); W=M+H
    A,first,leaf,count,offs=build_ref(keys)
    total=offs[n]
    val={}; nxt={}; par={}; arrive={}; gstart=[None]*n
    def quad(s,d): return (keys[s]>>(64-BITS*d))&3
    for b in range((n+M-1)//M):
        s0=b*M
        sA=[(A[s0+i] if s0+i<n else -1) for i in range(W)]
        sw=s0+W
        sA.append(A[sw] if sw<n else -1)
        soffs=[(offs[s0+i] if s0+i<=n else total) for i in range(W+1)]
        send=[None]*W; sval=[None]*W; pulled=[False]*W
        st=[]
        for i in range(W):
            s=s0+i
            cnt=count[s] if s<n else 0
            d=dict(cnt=cnt,i=i,s=s,alive=False)
            if cnt:
                t=s+1
                while t<n and keys[t]==keys[s]: t+=1
                d.update(firstd=first[s],leafd=leaf[s],off=offs[s],tp=min(t-s0,W+1),cur=('leaf',s),dcur=leaf[s])
                if i<M:
                    for dd in range(first[s],leaf[s]+1):
                        c=offs[s]+dd-first[s]
                        par[c]= c-1 if dd>first[s] else None
                        if dd==leaf[s]:
                            val[c]=('leaf',s); nxt[c]= c+1 if c+1<total else 0
            st.append(d)
        span_ok=lambda i,t: t<=W and t-i<=H
        maxd=0
        for d in st:
            if d['cnt']:
                maxd=max(maxd,d['leafd'])
                d['alive']=span_ok(d['i'],d['tp'])
                if d['alive'] and d['firstd']==d['leafd']:
                    sval[d['i']]=d['cur']; send[d['i']]=d['tp']
        for dep in range(maxd-1,-1,-1):
            pub=[]
            for d in st:
                if d['cnt'] and d['alive'] and dep>=d['firstd'] and dep<d['leafd']:
                    assert d['dcur']==dep+1
                    ch=[d['cur']]; t=d['tp']; ok=True
                    while t<=W and sA[t]==dep:
                        e = send[t] if t<W else None
                        if e is None: ok=False;break
                        ch.append(sval[t]); t=e
                    ok = ok and span_ok(d['i'],t)
                    if ok:
                        u=d['tp']
                        while u<t: pulled[u]=True; u=send[u]
                        d['cur']=tuple(ch); d['tp']=t; d['dcur']=dep
                        c=d['off']+dep-d['firstd']
                        if d['i']<M:
                            val[c]=d['cur']; nxt[c]= soffs[t] if s0+t<n else 0
                        if dep==d['firstd']: pub.append((d['i'],d['cur'],t))
                    else: d['alive']=False
            for i,v,t in pub: sval[i]=v; send[i]=t
        for d in st:
            if not (d['cnt'] and d['i']<M): continue
            s=d['s']; i=d['i']; firstd=d['firstd']; dcur=d['dcur']; off=d['off']
            start=None
            for dd in range(dcur,firstd,-1):
                c=off+dd-firstd; p=c-1
                a=arrive.setdefault(p,[0,set()]); a[0]+=1; a[1].add(quad(s,dd))
            if dcur>firstd: start=off+dcur-firstd
            if firstd>0 and not (dcur==firstd and pulled[i]):
                shp=64-BITS*(firstd-1); pp=keys[s]>>shp
                lo=s
                while lo>0 and (keys[lo-1]>>shp)>=pp: lo-=1
                parent_local=False
                if dcur==firstd and d['alive']:
                    t=d['tp']; ok=True
                    while t<=W and sA[t]==firstd-1:
                        e=send[t] if t<W else None
                        if e is None: ok=False;break
                        t=e
                    parent_local= ok and t<=W and (s0+t)-lo<=H
                if not parent_local:
                    p=offs[lo]+(firstd-1-first[lo])
                    par[off]=p
                    a=arrive.setdefault(p,[0,set()]); a[0]+=1; a[1].add(quad(s,firstd))
                    if dcur==firstd: start=off
            gstart[s]=start
    # climb
    slots={}
    arrived={}
    order=list(range(n)); random.shuffle(order)
    for s in order:
        if s>=n or count[s]==0 or gstart[s] is None: continue
        c=gstart[s]
        assert c in val, ('start cell has no value',s,c)
        v=val[c]; cells=(nxt[c] if nxt[c] else total)-c
        while True:
            p=par.get(c)
            if p is None: break
            q=None
            # find quadrant: need owner/depth of c
            a=arrive[p]
            if a[0]==1:
                pv=(v,); cells+=1
            else:
                slots.setdefault(p,[]).append((c,v,cells))
                arrived[p]=arrived.get(p,0)+1
                if arrived[p]!=a[0]: break
                chs=sorted(slots[p])   # cell index order == quadrant order
                pv=tuple(x[1] for x in chs); cells=1+sum(x[2] for x in chs)
            assert p not in val, ('parent computed twice',p)
            val[p]=pv; nxt[p]= p+cells if p+cells<total else 0
            v=pv; c=p
    return val,nxt,(A,first,leaf,count,offs)
def check(keys,M,H):
    keys=sorted(keys)
    val,nxt,(A,first,leaf,count,offs)=model(keys,M,H)
    rv,rn=ref_values(keys,A,first,leaf,count,offs)
    assert set(rv)==set(range(offs[-1]))
    for c in rv:
        assert c in val, ('missing',c)
        assert val[c]==rv[c], ('value',c)
        assert nxt[c]==rn[c], ('next',c,nxt[c],rn[c])


def first_with_prefix(keys, s, shp, pp):
    """mirror of bh_first_with_prefix: first j in [0, s] with (keys[j] >> shp) >= pp, 7 independent probes per round"""
    lo, hi = 0, s
    pos = [s - (1 << (3 * j)) if s >= (1 << (3 * j)) else 0 for j in range(7)]
    inr = [(keys[p] >> shp) >= pp for p in pos]
    found = False
    for j in range(7):
        if found:
            continue
        if inr[j]:
            hi = pos[j]
        else:
            lo = pos[j] + 1
            found = True
    while hi > lo:
        ln = hi - lo
        if ln <= 7:
            ans = hi
            for j in range(6, -1, -1):
                if j < ln and (keys[lo + j] >> shp) >= pp:
                    ans = lo + j
            return ans
        pos = [lo + (ln * (j + 1)) // 8 for j in range(7)]
        inr = [(keys[p] >> shp) >= pp for p in pos]
        nlo, nhi, closed = lo, hi, False
        for j in range(7):
            if closed:
                continue
            if inr[j]:
                nhi = pos[j]
                closed = True
            else:
                nlo = pos[j] + 1
        lo, hi = nlo, nhi
    return lo
