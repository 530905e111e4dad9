import sys, time, numpy as np, scipy.sparse as sp
sys.path.insert(0,'/root/repo')
from kakveda_b200 import synth
from kakveda_b200.similarity import Vocabulary
n=int(sys.argv[1]); nq=int(sys.argv[2]); K=16
CH=int(sys.argv[3]) if len(sys.argv)>3 else 64
t0=time.time()
v=Vocabulary()
buf,off=synth.signatures_packed(synth.CORPUS_SEED,0,n)
fb=v.featurize_packed(buf,off,0,grow=True)
ip=fb.indptr.copy(); ids=fb.ids.copy().astype(np.int64); tf=fb.tf.copy().astype(np.float64)
V=len(v)
rowof=np.repeat(np.arange(n),np.diff(ip))
df=np.bincount(ids,minlength=V).astype(np.float64)
idf_b=np.log((n+2)/(df+1))+1; idf_q=np.log((n+2)/(df+2))+1
a=idf_q**2; d=a-idf_b**2
B=np.bincount(rowof,weights=(tf*idf_b[ids])**2,minlength=n)
print("rows",n,"V",V,"nnz/row",len(ids)/n,"t",time.time()-t0)
qbuf,qoff=synth.signatures_packed(synth.QUERY_SEED,0,nq,dup_of_seed=synth.CORPUS_SEED,dup_rows=n)
qf=v.featurize_packed(qbuf,qoff,0,grow=False)
qip=qf.indptr.copy(); qids=qf.ids.copy().astype(np.int64); qtf=qf.tf.copy().astype(np.float64)
oov=np.array(qf.oov_tf2) if hasattr(qf,'oov_tf2') else np.zeros(nq)
print("query feats", len(qids)/nq, "oov attr", hasattr(qf,'oov_tf2'), [x for x in dir(qf) if not x.startswith('_')])
qrow=np.repeat(np.arange(nq),np.diff(qip))
idf0=np.log((n+2)/2.0)+1
nqv=np.bincount(qrow,weights=(qtf**2)*a[qids],minlength=nq)+oov*idf0**2
A=sp.csr_matrix((tf,ids,ip),shape=(n,V))
A2=sp.csr_matrix((tf*tf,ids,ip),shape=(n,V))
W=sp.csc_matrix((qtf*a[qids],(qids,qrow)),shape=(V,nq))
D=sp.csc_matrix((d[qids],(qids,qrow)),shape=(V,nq))
t0=time.time()
dot=np.asarray((A@W).todense()); corr=np.asarray((A2@D).todense())
score=dot/np.sqrt(nqv[None,:]*(B[:,None]+corr))
print("scores done",time.time()-t0)
part=np.partition(score,n-K,axis=0)[n-K]
theta=part
print("theta quantiles",np.quantile(theta,[0,.1,.25,.5,.75,.9,1]))
print("rows >= theta per query (mean/max)",(score>=theta[None,:]).sum(0).mean(),(score>=theta[None,:]).sum(0).max())
np.save('/tmp/sim/theta.npy',theta)
# per-feature classes
def chunk_eval(perm,label,CH=CH):
    pos_of=np.empty(n,dtype=np.int64); pos_of[perm]=np.arange(n)
    nch=(n+CH-1)//CH
    ch=pos_of[rowof]//CH
    # union with max tf
    U=sp.csr_matrix((tf,(ch,ids)),shape=(nch,V))  # sums duplicates; need max -> use tf==1 mostly; approximate with max via sort
    key=ch*V+ids
    o=np.lexsort((tf,key)); ks=key[o]; last=np.r_[ks[1:]!=ks[:-1],True]
    uc=ks[last]//V; uf=ks[last]%V; ut=tf[o][last]
    U=sp.csr_matrix((ut,(uc,uf)),shape=(nch,V)); U2=sp.csr_matrix((ut*ut,(uc,uf)),shape=(nch,V))
    minB=np.full(nch,np.inf); np.minimum.at(minB,pos_of//CH,B)
    bd=np.asarray((U@W).todense()); bc=np.asarray((U2@D).todense())
    den=nqv[None,:]*(minB[:,None]+bc)
    ub=np.where(den>0,bd/np.sqrt(np.maximum(den,1e-300)),np.inf)
    surv=ub>=theta[None,:]*0.99998
    print(label,"CH",CH,"chunks",nch,"entries/chunk-union %.1f"%(len(uc)/nch),"surviving (chunk,q) frac %.5f"%surv.mean(),"per query mean %.1f median %.1f max %d"%(surv.sum(0).mean(),np.median(surv.sum(0)),surv.sum(0).max()))
    # tile-level: queries sorted by text in tiles of 128
    return ub,surv
L=int(np.diff(ip).max())
def lexperm(idmat_cols,cls):
    return np.lexsort(idmat_cols[::-1]+[cls])
pad=np.zeros((n,L),dtype=np.int64); col=np.arange(len(ids))-np.repeat(ip[:-1],np.diff(ip)); pad[rowof,col]=ids+1
cls=np.floor(np.log2(B)*2).astype(np.int64)
perm1=np.lexsort([pad[:,j] for j in range(L-1,-1,-1)]+[cls])
ub1,s1=chunk_eval(perm1,"O1 token order + normclass")
# O2: features within row sorted by descending df
rank=np.empty(V,dtype=np.int64); rank[np.argsort(-df,kind='stable')]=np.arange(V)
o=np.lexsort((rank[ids],rowof)); ids2=ids[o]
pad2=np.full((n,L),V+1,dtype=np.int64); pad2[rowof,col]=rank[ids2]
perm2=np.lexsort([pad2[:,j] for j in range(L-1,-1,-1)]+[cls])
ub2,s2=chunk_eval(perm2,"O2 freq order + normclass")
perm3=np.lexsort([pad2[:,j] for j in range(L-1,-1,-1)])
ub3,s3=chunk_eval(perm3,"O3 freq order, no normclass")
# O4: rare-first order (ascending df) -> rows sharing rare words adjacent
pad4=np.full((n,L),-1,dtype=np.int64); o4=np.lexsort((-rank[ids],rowof)); pad4[rowof,col]=-rank[ids[o4]]
perm4=np.lexsort([pad4[:,j] for j in range(L-1,-1,-1)]+[cls])
ub4,s4=chunk_eval(perm4,"O4 rare-first order + normclass")
np.save('/tmp/sim/surv1.npy',s1)
