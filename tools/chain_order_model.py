"""Discrete-event model of the chained linear kernel tile list (gemm_chain.cu): 74 clusters walk the list round-robin;
a tile = loads (may start once its producers published and the previous tile's mainloop started) -> mainloop -> epilogue.
Compares the implemented phase-major order with diagonal wavefront orders.  Units: one K = 512 mainloop."""
import heapq, sys
C=74
def make(order_fn, mp=50):
    nblk=(2,4,2,6); Kc=(1,1,2,1); Ec=(1.3,0.9,1.3,0.6)   # mainloop / epilogue cost in units of a K=512 mainloop
    tiles=order_fn(mp,nblk)
    idx={t:i for i,t in enumerate(tiles)}
    return tiles, idx, nblk, Kc, Ec
def simulate(order_fn, lat=0.35, mp=50, verbose=False):
    tiles, idx, nblk, Kc, Ec = make(order_fn, mp)
    n=len(tiles)
    publish=[None]*n
    # per cluster state
    load_start=[0.0]*n; ml_start=[0.0]*n; ml_end=[0.0]*n; ep_end=[0.0]*n
    # process tiles in global index order: since a tile depends only on lower indices and on same-cluster predecessors (lower index), one pass works
    for i,(p,m,nb) in enumerate(tiles):
        c=i%C; prev=i-C; prev2=i-2*C
        dep=0.0
        if p>0:
            dep=max(publish[idx[(p-1,m,k)]] for k in range(nblk[p-1]))
        ls=max(dep, load_start[prev]+Kc[tiles[prev][0]]*0.0 if prev>=0 else 0.0, ml_start[prev] if prev>=0 else 0.0)  # prefetch once the previous tile's mainloop has started
        load_start[i]=ls
        ms=max(ls+lat, ml_end[prev] if prev>=0 else 0.0, ep_end[prev2] if prev2>=0 else 0.0)
        ml_start[i]=ms; ml_end[i]=ms+Kc[p]
        es=max(ml_end[i], ep_end[prev] if prev>=0 else 0.0)
        ep_end[i]=es+Ec[p]
        publish[i]=ep_end[i]+0.15
    total=max(ep_end)
    work=sum(Kc[p] for p,_,_ in tiles)/C
    return total, work
def phase_major(mp,nblk):
    return [(p,m,n) for p in range(4) for m in range(mp) for n in range(nblk[p])]
def diagonal(D):
    def f(mp,nblk):
        out=[]
        for w in range(mp+3*D):
            for p in range(4):
                m=w-p*D
                if 0<=m<mp:
                    out+=[(p,m,n) for n in range(nblk[p])]
        return out
    return f
def diag2(D01,D12,D23):
    def f(mp,nblk):
        out=[]
        offs=[0,D01,D01+D12,D01+D12+D23]
        for w in range(mp+offs[3]):
            for p in range(4):
                m=w-offs[p]
                if 0<=m<mp: out+=[(p,m,n) for n in range(nblk[p])]
        return out
    return f
print("phase-major", simulate(phase_major))
for D in (4,8,12,16,20,25,30,40):
    print("diag",D, simulate(diagonal(D)))
for a,b,c in ((40,12,12),(45,10,10),(50,10,5),(50,15,10),(50,25,12),(30,20,10)):
    print("diag2",a,b,c, simulate(diag2(a,b,c)))
