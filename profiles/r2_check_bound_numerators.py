import sys, ctypes as C, numpy as np
sys.path.insert(0,'.')
from kakveda_b200 import synth, GfkbIndex, _capi
n,q=int(sys.argv[1]),int(sys.argv[2])
ix=GfkbIndex()
buf,off=synth.signatures_packed(synth.CORPUS_SEED,0,n)
fb=ix.vocab.featurize_packed(buf,off,0,grow=True)
ip=fb.indptr.copy(); ids=fb.ids.copy().astype(np.int64); tf=fb.tf.copy().astype(np.float64)
ix.add_features(fb); fb.close(); ix.finalize()
V=len(ix.vocab)
qbuf,qoff=synth.signatures_packed(synth.QUERY_SEED,0,q,dup_of_seed=synth.CORPUS_SEED,dup_rows=n)
qfb=ix.vocab.featurize_packed(qbuf,qoff,0,grow=False)
qip=qfb.indptr.copy(); qids=qfb.ids.copy().astype(np.int64); qtf=qfb.tf.copy().astype(np.float64)
ix.upload_queries(qfb)
nch=(n+31)//32
out=np.zeros((q,nch),dtype=np.float32); sq=np.zeros(q,dtype=np.int32)
_capi.check(_capi.load().kv_debug_bound_numerators(ix._h,16,out.ctypes.data_as(C.POINTER(C.c_float)),sq.ctypes.data_as(C.POINTER(C.c_int32))))
# numpy union bound
import scipy.sparse as sp
rowof=np.repeat(np.arange(n),np.diff(ip))
df=np.bincount(ids,minlength=V).astype(np.float64)
idf_q=np.log((n+2)/(df+2))+1; a=idf_q**2
idf_b=np.log((n+2)/(df+1))+1
B=np.bincount(rowof,weights=(tf*idf_b[ids])**2,minlength=n)
L=int(np.diff(ip).max()); col=np.arange(len(ids))-np.repeat(ip[:-1],np.diff(ip))
pad=np.zeros((n,L),dtype=np.int64); pad[rowof,col]=ids+1
B32=B.astype(np.float32)
cls=np.where(B32>0,np.floor(np.log2(B32.astype(np.float64))*2),-1000).astype(np.int64)
perm=np.lexsort([pad[:,j] for j in range(L-1,-1,-1)]+[cls])
pos_of=np.empty(n,dtype=np.int64); pos_of[perm]=np.arange(n)
ch=pos_of[rowof]//32
key=ch*V+ids; o=np.lexsort((tf,key)); ks=key[o]; last=np.r_[ks[1:]!=ks[:-1],True]
U=sp.csr_matrix((tf[o][last],(ks[last]//V,ks[last]%V)),shape=(nch,V))
qrow=np.repeat(np.arange(q),np.diff(qip)); known=qids<V
W=sp.csc_matrix((qtf[known]*a[qids[known]],(qids[known],qrow[known])),shape=(V,q))
ub=np.asarray((U@W).todense()).T   # [q][nch]
got=out  # by slot
want=ub[sq]
ratio=(got+1e-3)/(want+1e-3)
print("slots",q,"chunks",nch,"min ratio %.5f"%ratio.min(),"median %.4f"%np.median(ratio),"mean %.4f"%ratio.mean(),"p99 %.4f"%np.quantile(ratio,0.99),"max %.3f"%ratio.max())
bad=np.argwhere(ratio<0.9999)
print("too low:",len(bad), bad[:5])
hi=np.argwhere(ratio>1.05); print("above 1.05:",len(hi),"of",ratio.size)
if len(hi):
    i,j=hi[0]; print("example",i,j,got[i,j],want[i,j])
    d=got-want; print("excess quantiles",np.quantile(d,[.5,.9,.99,.999,1]))
d=got-want
big=np.argwhere(d>20)
print("pairs with excess > 20:",len(big))
for (i,j) in big[:6]:
    b0=(j//128)*128
    blk=d[i,b0:b0+128]
    print("slot",i,"chunk",j,"excess %.1f"%d[i,j],"| same block: n>20:",int((blk>20).sum()),"min %.2f median %.2f"%(blk.min(),np.median(blk)), "| same 32-col group n>20:", int((d[i,(j//32)*32:(j//32)*32+32]>20).sum()))
# how are the big excesses distributed over chunks (columns)?
cols=np.bincount(big[:,1]%128,minlength=128); print("by column-in-block (first 16):",cols[:16], "max col", cols.argmax(), cols.max())
rows=np.bincount(big[:,0],minlength=q); print("slots with most:",np.argsort(-rows)[:5], np.sort(rows)[::-1][:5])
