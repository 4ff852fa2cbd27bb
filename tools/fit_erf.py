"""Weighted minimax fit of erf(t) = 1 - 2^(t P(t)) on [0, 4] used by the fused GELU epilogue (csrc/al_tc.cuh gelu_erf)."""
import numpy as np
from scipy.special import erfc, erf
from numpy.polynomial import chebyshev as C
T=4.0
t=np.cos(np.pi*(np.arange(4000)+0.5)/4000)*T/2+T/2
t=np.sort(t)
q=np.log2(erfc(t))/t
for deg in (5,6,7,8):
    w=erfc(t)*t*np.log(2)
    # iterative reweighting towards minimax
    ww=w.copy()
    for it in range(60):
        V=np.vander(t,deg+1,increasing=True)
        coef,*_=np.linalg.lstsq(V*ww[:,None],q*ww,rcond=None)
        approx=1-np.exp2(t*(V@coef))
        err=np.abs(approx-erf(t))
        ww=ww*(1+ 2*err/err.max())
        ww/=ww.max()
    # float32 eval
    c32=coef.astype(np.float32)
    tt=np.linspace(0,6,200001).astype(np.float32)
    tc=np.minimum(tt,np.float32(T))
    p=np.zeros_like(tc)+c32[-1]
    for c in c32[-2::-1]:
        p=p*tc+c
    a=(np.float32(1)-np.exp2(tc*p)).astype(np.float32)
    e=np.abs(a.astype(np.float64)-erf(tt.astype(np.float64)))
    print(deg, err.max(), e.max(), [float(x) for x in c32])
