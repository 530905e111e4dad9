"""CPU measurement (no GPU) of how the hit events of K1b's bound pass are distributed over features.

    python profiles/sim_bound_pass_hits.py ROWS QUERIES        # e.g. 1000000 4096

Builds the real host layout (C++ featuriser, (norm class, text) order, 64-row chunk unions, universal and
summary-universal folding), forms tiles of 128 text-sorted queries and counts, per tile, the hit events one pass over
all chunk summaries produces (sum over the tile's table features of the number of chunk unions containing the
feature), and which share of them falls on the most frequent features.  Result at 1M rows x 4096 queries:
85.5 hits per chunk summary and tile (ncu-derived estimate on the GPU: ~100); a GLOBAL set of the 256 most
chunk-frequent features covers 82 % of the hits (512: 92 %; per-tile top-256: 93 %).  This is the basis of the round-2
plan in DESIGN.md section 8: evaluate those 256 features densely (small fp16 tensor-core products with weights rounded
up, valid upper bounds) and leave only the rare features (~15 hits per chunk) to the event loop.
"""
import sys, time, numpy as np, scipy.sparse as sp
sys.path.insert(0,'/root/repo')
from kakveda_b200 import synth
from kakveda_b200.similarity import Vocabulary
n=int(sys.argv[1]); nq_total=int(sys.argv[2])
v=Vocabulary()
buf,off=synth.signatures_packed(synth.CORPUS_SEED,0,n)
fb=v.featurize_packed(buf,off,0,grow=True)
ip=fb.indptr.copy(); ids=fb.ids.copy().astype(np.int64); tf=fb.tf.copy().astype(np.float64)
V=len(v)
rowof=np.repeat(np.arange(n),np.diff(ip))
df=np.bincount(ids,minlength=V).astype(np.float64)
idf_b=np.log((n+2)/(df+1))+1
B=np.bincount(rowof,weights=(tf*idf_b[ids])**2,minlength=n)
univ=(df==n)
L=int(np.diff(ip).max()); pad=np.zeros((n,L),dtype=np.int64)
col=np.arange(len(ids))-np.repeat(ip[:-1],np.diff(ip)); pad[rowof,col]=ids+1
cls=np.floor(np.log2(B)*2).astype(np.int64)
perm=np.lexsort([pad[:,j] for j in range(L-1,-1,-1)]+[cls])
pos_of=np.empty(n,dtype=np.int64); pos_of[perm]=np.arange(n)
nch=(n+63)//64
nu=~univ[ids]
key=np.unique((pos_of[rowof]//64)[nu]*V+ids[nu])
sc=key//V; sf=key%V
chunkfreq=np.bincount(sf,minlength=V)
su=chunkfreq>=0.9*nch
print("chunks",nch,"summary entries/chunk",len(key)/nch,"summary-universal",su.sum())
# queries
qbuf,qoff=synth.signatures_packed(synth.QUERY_SEED,0,nq_total,dup_of_seed=synth.CORPUS_SEED,dup_rows=n)
qf=v.featurize_packed(qbuf,qoff,0,grow=False)
qip=qf.indptr.copy(); qids=qf.ids.copy().astype(np.int64)
Lq=int(np.diff(qip).max()); qpad=np.zeros((nq_total,Lq),dtype=np.int64)
qrow=np.repeat(np.arange(nq_total),np.diff(qip)); qcol=np.arange(len(qids))-np.repeat(qip[:-1],np.diff(qip))
qpad[qrow,qcol]=qids+1
qperm=np.lexsort([qpad[:,j] for j in range(Lq-1,-1,-1)])
tot=0; res=[]
for t in range(nq_total//128):
    qs=qperm[t*128:(t+1)*128]
    feats=np.unique(np.concatenate([qids[qip[q]:qip[q+1]] for q in qs]))
    feats=feats[(feats<V)]
    feats=feats[~univ[feats] & ~su[feats]]
    cf=np.sort(chunkfreq[feats])[::-1]
    hits=cf.sum()
    res.append((len(feats),hits/nch,cf[:64].sum()/hits,cf[:128].sum()/hits,cf[:256].sum()/hits))
r=np.array(res)
print("tiles",len(r),"table feats %.0f hits/chunk %.1f share top64 %.3f top128 %.3f top256 %.3f"%tuple(r.mean(axis=0)))
# global D: top features by chunkfreq (not tile specific)
order=np.argsort(-np.where(univ|su,0,chunkfreq))
for D in (128,256,512):
    inD=np.zeros(V,bool); inD[order[:D]]=True
    sh=[]
    for t in range(nq_total//128):
        qs=qperm[t*128:(t+1)*128]
        feats=np.unique(np.concatenate([qids[qip[q]:qip[q+1]] for q in qs])); feats=feats[feats<V]; feats=feats[~univ[feats]&~su[feats]]
        sh.append(chunkfreq[feats[inD[feats]]].sum()/max(1,chunkfreq[feats].sum()))
    print("global D",D,"share of hits covered %.3f"%np.mean(sh), "min chunkfreq in D", chunkfreq[order[D-1]]/nch)
