"""The exact-walk kernel (libcimbar_b200/csrc/k1x_flood.cu, heap_pop) resolves five heap levels of libstdc++'s
__adjust_heap per memory round trip: the 31 lanes of a warp load the child pairs of the five-level subtree under the hole,
one ballot collects "left child preferred", and every lane decides from its ancestors' bits whether the descent passes
through its node.  Nothing is stored during the descent: the final position of the sifted-up last element is decided from
the values the path lanes hold (s = 1 + deepest level whose moved value may stay), then s + 1 stores are made.  This is a lane-by-lane model of that routine (same per-lane constants, same formulas) checked against
the literal libstdc++ push_heap / pop_heap on random sequences with many equal priorities -- the pop order of ties is what
FloodDecodePositions' results depend on (src/lib/cimb_translator/FloodDecodePositions.cpp:49-67).  Host logic, no GPU."""
import random


def prio(e): return e>>25
# literal libstdc++
def ref_push(v,e):
    v.append(e); hole=len(v)-1
    while hole>0:
        p=(hole-1)>>1
        if prio(v[p])<=prio(e): break
        v[hole]=v[p]; hole=p
    v[hole]=e
def ref_pop(v):
    top=v[0]; value=v[-1]; v.pop(); ln=len(v)
    if ln==0: return top
    hole=0; second=0
    while second < (ln-1)//2:
        second=2*(second+1)
        if prio(v[second])>prio(v[second-1]): second-=1
        v[hole]=v[second]; hole=second
    if (ln&1)==0 and second==(ln-2)//2:
        second=2*(second+1); v[hole]=v[second-1]; hole=second-1
    while hole>0:
        p=(hole-1)>>1
        if prio(v[p])<=prio(value): break
        v[hole]=v[p]; hole=p
    v[hole]=value
    return top
# lane-parallel emulation
def lanes():
    out=[]
    for lane in range(32):
        d=(lane+1).bit_length()-1; j=lane+1-(1<<d); dp=d-1 if d>0 else 0
        am=aw=0; i=lane
        while i>0:
            p=(i-1)>>1; am|=1<<p
            if i&1: aw|=1<<p
            i=p
        valid=lane<31
        if lane==31: d=0;j=0;dp=0;am=0;aw=0
        out.append((d,j,dp,am,aw,valid))
    return out
SL=lanes()
def par_pop(v):
    top=v[0]; value=v[-1]; v.pop(); ln=len(v)
    if ln==0: return top
    lim=(ln-1)>>1; hole=0; level=0
    saved=[]   # per round: per lane (lvl,node,m)
    for r in range(3):          # kPopRounds
        rec=[(-1,0,0)]*32
        if hole<lim:
            st=[]; pref=0
            for lane in range(32):
                d,j,dp,am,aw,valid=SL[lane]
                node=((hole+1)<<d)-1+j
                has2=valid and node<lim
                c=(v[2*node+1],v[2*node+2]) if has2 else (0,0)
                left=prio(c[1])>prio(c[0])
                if has2 and left: pref|=1<<lane
                st.append((node,has2,c,left))
            eb=0; nxts=[]; lv=[]
            rec=[]
            for lane in range(32):
                d,j,dp,am,aw,valid=SL[lane]
                node,has2,c,left=st[lane]
                parent_ok=(d==0) or (((hole+1)<<dp)-1+(j>>1))<lim
                reached=valid and parent_ok and (((pref^aw)&am)==0)
                rec.append((level+d,node,c[0] if left else c[1]) if (reached and has2) else (-1,0,0))
                ends=reached and ((not has2) or d==4)
                if ends: eb|=1<<lane
                nxts.append(2*node+2-(1 if left else 0) if has2 else node)
                lv.append(d+(1 if has2 else 0))
            assert bin(eb).count("1")==1
            src=(eb&-eb).bit_length()-1
            hole=nxts[src]; level+=lv[src]
        saved.append(rec)
    lone=(ln&1)==0 and hole==(ln-2)>>1
    node_lone=hole; m_lone=0
    if lone: m_lone=v[2*hole+1]; hole=2*hole+1
    lvl_lone=level
    vp=prio(value); kmax=-1
    for rec in saved:
        for (lvl,node,m) in rec:
            if lvl>=0 and prio(m)<=vp: kmax=max(kmax,lvl)
    if lone and prio(m_lone)<=vp: kmax=lvl_lone
    s=kmax+1; last_level=level+(1 if lone else 0)
    for rec in saved:
        for (lvl,node,m) in rec:
            if lvl>=0 and lvl<=s: v[node]=value if lvl==s else m
    if lone and lvl_lone<=s: v[node_lone]=value if lvl_lone==s else m_lone
    if s==last_level: v[hole]=value
    return top


def test_lane_parallel_pop_equals_libstdcxx_pop():
    random.seed(1)
    for trial in range(40):
        a, b, uid = [], [], 0
        maxp = random.choice([1, 2, 3, 8, 64])
        for step in range(random.choice([50, 500, 5000, 30000])):
            if not a or random.random() < random.choice([0.5, 0.6, 0.8]):
                e = (random.randrange(maxp) << 25) | uid
                uid += 1
                ref_push(a, e)
                ref_push(b, e)
            else:
                assert ref_pop(a) == par_pop(b)
            assert a == b, (trial, step)
        while a:
            assert ref_pop(a) == par_pop(b) and a == b
