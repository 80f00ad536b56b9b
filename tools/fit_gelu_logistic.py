"""Minimax (LP) fit of the logistic-form exact-erf GELU used by gelu_fast2 (csrc/sdes_common.cuh):
GELU(x) = x / (1 + 2^(-x P(x^2))); prints the coefficients of P per degree with float64 / emulated-fp32 errors."""
import numpy as np
from scipy.special import erfc, log_ndtr
from scipy.optimize import linprog
def gelu(x): return x*0.5*erfc(-x/np.sqrt(2))
X=6.5
x=np.concatenate([np.linspace(1e-4,X,4001)])
q=(log_ndtr(x)-log_ndtr(-x))/np.log(2)
u=x*x
sig=1/(1+2.0**(-q)); dsig=sig*(1-sig)*np.log(2)
wt=x*x*dsig   # d gelu / d P
for deg in (3,4,5,6):
    V=np.vander(u,deg+1,increasing=True)
    # minimize t s.t. |wt*(V c - q/x)| <= t
    n=deg+1
    A=np.block([[ (wt[:,None]*V), -np.ones((len(x),1))],[-(wt[:,None]*V), -np.ones((len(x),1))]])
    b=np.concatenate([wt*q/x, -wt*q/x])
    cost=np.zeros(n+1); cost[-1]=1
    res=linprog(cost,A_ub=A,b_ub=b,bounds=[(None,None)]*(n+1),method='highs')
    c=res.x[:n]
    xx=np.linspace(-9,9,2000001)
    def f64(xx):
        uu=xx*xx; p=np.polyval(c[::-1],uu); return xx/(1+2.0**(-xx*p))
    e64=np.abs(f64(xx)-gelu(xx)).max()
    # fp32 emulation
    x32=xx.astype(np.float32); u32=x32*x32
    p=np.float32(c[-1])*np.ones_like(x32)
    for k in range(deg-1,-1,-1): p=(p*u32+np.float32(c[k])).astype(np.float32)
    qq=(x32*p).astype(np.float32)
    with np.errstate(over='ignore'):
        e=np.exp2(-qq.astype(np.float64)).astype(np.float32)
    d=(np.float32(1)+e).astype(np.float32)
    out=(x32*(np.float32(1)/d)).astype(np.float32)
    e32=np.abs(out.astype(np.float64)-gelu(x32.astype(np.float64))).max()
    print(deg,'lin-minimax t=%.3g'%res.x[-1],'f64 err %.3g'%e64,'f32 err %.3g'%e32, repr(c))
