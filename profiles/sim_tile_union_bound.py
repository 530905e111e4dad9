"""CPU simulation (NumPy/SciPy, no GPU) of a design idea for K1b's bound pass that was REJECTED in round 1.

Idea: replace the per-query evaluation of every chunk summary (47 % of the kernel's cycles) by ONE evaluation per
(tile, chunk) against a pseudo-query -- the union of the tile's features with the largest weight, the smallest |q|^2
and the most negative correction -- which is a valid upper bound of every query's bound, costs ~3 % of the per-query
pass, and would leave the exact per-query bound to the chunks that survive it.

    python profiles/sim_tile_union_bound.py ROWS QUERIES TILES      # e.g. 300000 4096 6

builds the real layout on the host (C++ featuriser, (norm class, text) order, 64-row chunk unions), takes TILES tiles of
128 text-sorted queries, computes their true top-16 thresholds from exact float64 scores, and prints per tile the
fraction of chunks whose bound reaches the tile's weakest threshold: exact tile-max bound (what K1b stores today),
pseudo-query over 128 / 32 / 16 queries, and the fraction some query really needs.  Result (300k rows, 4096 queries):
exact 0.62, pseudo-query 1.000 / 1.000 / 1.000, needed 0.38 -- the pseudo-query bound never prunes anything (the
smallest norm combined with the union of all rare words overshoots), so the idea was dropped.  Kept as the harness for
evaluating bound-pass variants (chunk size, grouping, orderings) without GPU time.
"""
import sys, time, numpy as np, scipy.sparse as sp
sys.path.insert(0,'/root/repo')
from kakveda_b200 import synth
from kakveda_b200.similarity import Vocabulary
n=int(sys.argv[1]); nq_total=int(sys.argv[2]); ntiles=int(sys.argv[3])
t0=time.time()
v=Vocabulary()
buf,off=synth.signatures_packed(synth.CORPUS_SEED,0,n)
fb=v.featurize_packed(buf,off,0,grow=True)
ip=fb.indptr.copy(); ids=fb.ids.copy().astype(np.int64); tf=fb.tf.copy().astype(np.float64)
V=len(v); print("rows",n,"V",V,"nnz",len(ids), time.time()-t0)
rowof=np.repeat(np.arange(n),np.diff(ip))
C=sp.csr_matrix((tf,ids,ip),shape=(n,V))
df=np.bincount(ids,minlength=V).astype(np.float64)
idf_b=np.log((n+2)/(df+1))+1; idf_q=np.log((n+2)/(df+2))+1
a=idf_q**2; d=a-idf_b**2
univ=(df==n)
B=np.asarray(C.multiply(C)@(idf_b**2)).ravel()
# sort rows by (norm class, id sequence)
L=int(np.diff(ip).max())
pad=np.zeros((n,L),dtype=np.int64)
lens=np.diff(ip)
col=np.arange(len(ids))-np.repeat(ip[:-1],lens)
pad[rowof,col]=ids+1
cls=np.floor(np.log2(B)*2).astype(np.int64)
keys=[pad[:,j] for j in range(L-1,-1,-1)]+[cls]
perm=np.lexsort(keys)
print("sorted",time.time()-t0)
pos_of=np.empty(n,dtype=np.int64); pos_of[perm]=np.arange(n)
chunk_of_entry=pos_of[rowof]//64
nch=(n+63)//64
nu=~univ[ids]
key=chunk_of_entry[nu]*V+ids[nu]
order=np.argsort(key,kind='stable')
ks=key[order]; tfs=tf[nu][order]
first=np.concatenate([[True],ks[1:]!=ks[:-1]])
grp=np.cumsum(first)-1
maxtf=np.zeros(grp[-1]+1); np.maximum.at(maxtf,grp,tfs)
uk=ks[first]; sc=uk//V; sf=uk%V
S=sp.csr_matrix((maxtf,(sc,sf)),shape=(nch,V))
print("summaries",S.nnz/nch,"entries/chunk",time.time()-t0)
Bpos=B[perm]
Bmin=np.minimum.reduceat(Bpos,np.arange(0,n,64))
# queries
qbuf,qoff=synth.signatures_packed(synth.QUERY_SEED,0,nq_total,dup_of_seed=synth.CORPUS_SEED,dup_rows=n)
qf=v.featurize_packed(qbuf,qoff,0,grow=False)
qip=qf.indptr.copy(); qids=qf.ids.copy().astype(np.int64); qtf=qf.tf.copy().astype(np.float64); oov=qf.oov.copy()
idf0=np.log((n+2)/2)+1
inv=qids<V
Qall=sp.csr_matrix((qtf[inv]*a[qids[inv]], qids[inv], np.concatenate([[0],np.cumsum(np.add.reduceat(inv.astype(np.int64),qip[:-1]))]) ),shape=(nq_total,V)) if inv.all() else None
assert inv.all()
nqv=np.asarray(sp.csr_matrix((qtf**2*a[qids],qids,qip),shape=(nq_total,V)).sum(axis=1)).ravel()+oov*idf0**2
Lq=int(np.diff(qip).max()); qpad=np.zeros((nq_total,Lq),dtype=np.int64)
qrow=np.repeat(np.arange(nq_total),np.diff(qip)); qcol=np.arange(len(qids))-np.repeat(qip[:-1],np.diff(qip))
qpad[qrow,qcol]=qids+1
qcls=np.floor(np.log2(nqv)*2).astype(np.int64)
qperm=np.lexsort([qpad[:,j] for j in range(Lq-1,-1,-1)]+[qcls])
Qw=sp.csr_matrix((qtf*a[qids],qids,qip),shape=(nq_total,V))[qperm]
Qm=sp.csr_matrix((np.ones(len(qids)),qids,qip),shape=(nq_total,V))[qperm]
nqs=nqv[qperm]
Cd=C.multiply(C)@sp.diags(d)  # rows: tf^2 d
Sd=S.multiply(S)@sp.diags(d)
ucols=np.where(univ)[0]
print("universal",len(ucols))
res=[]
tiles=np.linspace(0,nq_total//128-1,ntiles).astype(int)
for t in tiles:
    sl=slice(t*128,(t+1)*128)
    qw=Qw[sl]; qm=Qm[sl]; nq_=nqs[sl]
    # exact scores vs all rows -> thresholds
    dot=(qw@C.T).toarray(); corr=(qm@Cd.T).toarray()
    den=nq_[:,None]*(B[None,:]+corr)
    sc_=np.where(den>0,dot/np.sqrt(np.maximum(den,1e-300)),0)
    thr=np.partition(sc_,n-16,axis=1)[:,n-16]
    thr_min=thr.min()
    # per-query chunk bounds (universal features are inside S? no: S excludes universal) -> add universal parts
    qwu=np.asarray(qw[:,ucols].sum(axis=1)).ravel()  # tf_c of universal = 1 assumed
    qcu=np.asarray((qm[:,ucols]@sp.diags(d[ucols])).sum(axis=1)).ravel()
    bd=(qw@S.T).toarray()+qwu[:,None]; bc=(qm@Sd.T).toarray()+qcu[:,None]
    bden=nq_[:,None]*(Bmin[None,:]+bc)
    bq=np.where(bden>0,bd/np.sqrt(np.maximum(bden,1e-300)),np.inf)
    tilemax=bq.max(axis=0)
    # loose: one pseudo query
    def loose(idx):
        wmax=np.asarray(qw[idx].max(axis=0).todense()).ravel(); anym=(np.asarray(qm[idx].sum(axis=0)).ravel()>0).astype(np.float64)
        dm=S@wmax+qwu[idx].max(); cm=Sd@anym+qcu[idx].min()
        den=nq_[idx].min()*(Bmin+cm)
        return np.where(den>0,dm/np.sqrt(np.maximum(den,1e-300)),np.inf)
    l1=loose(np.arange(128))
    l4=np.max([loose(np.arange(g*32,(g+1)*32)) for g in range(4)],axis=0)
    l8=np.max([loose(np.arange(g*16,(g+1)*16)) for g in range(8)],axis=0)
    # per-query requery pass (chunk must be scanned for some q)
    need=(bq>=thr[:,None]).any(axis=0)
    f=lambda x:(x>=thr_min).mean()
    res.append((t,thr_min,f(tilemax),f(l1),f(l4),f(l8),need.mean()))
    print("tile",t,"thr_min %.3f"%thr_min,"exact %.3f loose1 %.3f loose4 %.3f loose8 %.3f need %.3f"%tuple(res[-1][2:]),time.time()-t0,flush=True)
r=np.array(res)
print("mean exact %.3f loose1 %.3f loose4 %.3f loose8 %.3f need %.3f"%tuple(r[:,2:].mean(axis=0)))
