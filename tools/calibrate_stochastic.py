import numpy as np, math
S=np.load('/tmp/fmr_pool.npy'); P=S[:,:128]; V=S[:,128]
pg=P.mean(0); kg=int(np.argmax(pg[1:]))+1; lo,hi=max(kg-6,1),kg+7
print("kg",kg,"lo,hi",lo,hi)
k=np.arange(lo,hi)
band=P[:,lo:hi].sum(1); cen=(P[:,lo:hi]*k).sum(1)/band
print("centroid mean/std",cen.mean(),cen.std(),"band mean/std",band.mean(),band.std(), "var mean/std", V.mean(), V.std())
from scipy import stats
for name,x in (("cen",cen),("band",band),("var",V)):
    print(name,"skew",stats.skew(x),"kurt",stats.kurtosis(x))
rng=np.random.default_rng(0)
ng,nr=4096,1024
R=400000
zs={n:np.zeros(R) for n in ("cen","band","var")}
N=len(cen)
for r in range(R):
    ig=rng.integers(0,N,ng); ir=rng.integers(0,N,nr)
    for name,x in (("cen",cen),("band",band),("var",V)):
        a,b=x[ig],x[ir]
        if name=="cen":
            pooled=(a.var()*ng+b.var()*nr)/(ng+nr)
            z=(a.mean()-b.mean())/math.sqrt(pooled*(1/ng+1/nr))
        else:
            pooled=((a/a.mean()).var()*ng+(b/b.mean()).var()*nr)/(ng+nr)
            z=(a.mean()/b.mean()-1)/math.sqrt(pooled*(1/ng+1/nr))
        zs[name][r]=z
for name in zs:
    z=zs[name]
    print(name,"std",z.std(),"max|z|",np.abs(z).max(),"P(|z|>4)",(np.abs(z)>4).mean(),"P(|z|>4.5)",(np.abs(z)>4.5).mean(), "expected P>4", 6.3e-5)
