"""Protocol simulator of the PERSISTENT variant of the pair kernel (csrc/attention_tc2.cu; built, measured and not kept —
profiles/r02_notes_measured_dead_ends.md — the shipped kernel runs one item per CTA): every mbarrier with its arrival count and
phase, every warp role as a coroutine that waits on phase PARITIES exactly like the kernel does (TMA loads and MMA
commits complete at once).  Run on the CPU: prints ok / DEADLOCK for item streams with and without a second query tile and
1..5 key tiles per item.  It found the one deadlock of the first persistent version (group 1 racing through the items it
sits out and completing validity-word phases alone) before a second GPU run was spent on it."""
import sys
class Bar:
    def __init__(s, count): s.count=count; s.pending=count; s.phase=0
    def arrive(s):
        s.pending-=1
        if s.pending==0: s.pending=s.count; s.phase+=1
    def test(s, parity): return (s.phase & 1) != parity
def run(n_kv, n_items_list_has1):
    B={k:[Bar(c),Bar(c)] for k,c in dict(q=1,qfree=1,k=1,v=1,kfree=1,vfree=1,s=1,p=4,pv=1,ofree=4,vw=1,vwfree=8).items()}
    items=n_items_list_has1
    log=[]
    def producer():
        n0=n1=kt=0
        for has1 in items:
            def load_kv(j,kv):
                g=kt+j; s=g&1; use=g>>1
                fb=B['kfree' if kv else 'vfree'][s]
                if g>=2:
                    while not fb.test((use-1)&1): yield ('prod wait free',kv,g)
                B['k' if kv else 'v'][s].arrive()
            if n0>0:
                while not B['qfree'][0].test((n0-1)&1): yield ('prod wait qfree0',n0)
            B['q'][0].arrive()
            yield from load_kv(0,True)
            if has1:
                if n1>0:
                    while not B['qfree'][1].test((n1-1)&1): yield ('prod wait qfree1',n1)
                B['q'][1].arrive()
            yield from load_kv(0,False)
            for j in range(1,n_kv):
                yield from load_kv(j,True); yield from load_kv(j,False)
            n0+=1; n1+= 1 if has1 else 0; kt+=n_kv
    def mma():
        n0=n1=kt=0; c=[0,0]
        for has1 in items:
            def issue_s(w,j):
                B['s'][w].arrive()
                if j==n_kv-1: B['qfree'][w].arrive()
            def wait_k(j):
                g=kt+j
                while not B['k'][g&1].test((g>>1)&1): yield ('mma wait k',g)
            while not B['q'][0].test(n0&1): yield ('mma wait q0',n0)
            yield from wait_k(0)
            issue_s(0,0)
            if has1:
                while not B['q'][1].test(n1&1): yield ('mma wait q1',n1)
                issue_s(1,0)
            B['kfree'][kt&1].arrive()
            for j in range(n_kv):
                g=kt+j; s=g&1; more=j+1<n_kv
                while not B['v'][s].test((g>>1)&1): yield ('mma wait v',g)
                while not B['p'][0].test(c[0]&1): yield ('mma wait p0',c[0])
                if j==0 and n0>0:
                    while not B['ofree'][0].test((n0-1)&1): yield ('mma wait ofree0',n0)
                B['pv'][0].arrive(); c[0]+=1
                if more:
                    yield from wait_k(j+1)
                    issue_s(0,j+1)
                if has1:
                    while not B['p'][1].test(c[1]&1): yield ('mma wait p1',c[1])
                    if j==0 and n1>0:
                        while not B['ofree'][1].test((n1-1)&1): yield ('mma wait ofree1',n1)
                    B['pv'][1].arrive(); c[1]+=1
                B['vfree'][s].arrive()
                if more:
                    if has1: issue_s(1,j+1)
                    B['kfree'][s^1].arrive()
            n0+=1; n1+=1 if has1 else 0; kt+=n_kv
    def stager():
        for n,has1 in enumerate(items):
            buf=n&1
            if n>=2:
                while not B['vwfree'][buf].test(((n>>1)-1)&1): yield ('stager wait',n)
            B['vw'][buf].arrive()
    def wg(w, warp):
        cw=0
        for n,has1 in enumerate(items):
            buf=n&1
            while not B['vw'][buf].test((n>>1)&1): yield ('wg wait vw',w,n)
            if w==0 or has1:
                for j in range(n_kv):
                    while not B['s'][w].test(cw&1): yield ('wg wait s',w,cw)
                    B['p'][w].arrive(); cw+=1
                while not B['pv'][w].test((cw-1)&1): yield ('wg wait pv',w,cw)
                B['ofree'][w].arrive()
            B['vwfree'][buf].arrive()
    threads=[('prod',producer()),('mma',mma()),('stager',stager())]+[(f'wg{w}.{i}',wg(w,i)) for w in (0,1) for i in range(4)]
    state={}
    alive=dict(threads)
    for it in range(100000):
        progressed=False
        for name,g in list(alive.items()):
            prev=state.get(name)
            try:
                st=next(g)
                if st!=prev: progressed=True
                state[name]=st
            except StopIteration:
                del alive[name]; progressed=True
        if not alive: return 'ok'
        if not progressed:
            return 'DEADLOCK '+str(state)
    return 'timeout'
for n_kv in (1,2,3,4,5):
    for pat in ([True]*6,[False]*6,[True,False]*3):
        print(n_kv, pat[:2], run(n_kv,pat))
