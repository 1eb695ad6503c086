import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from smelter_b200.api import Context, Image, run_elementwise
ctx = Context(0)
rng = np.random.default_rng(0)
def img(shape):
    return Image.fromArray(ctx, rng.integers(0, 0x3C00, size=shape, dtype=np.uint16).view(np.float16))
S=(64,256,56,56)
def unary(kind):
    x=img(S); y,ms=run_elementwise(ctx,"unary",x,out_shape=S,iters=10,sub=kind); print("unary",kind,round(ms*1e3,1),"us",flush=True)
def binary(iters=10):
    x=img(S); x2=img(S); y,ms=run_elementwise(ctx,"binary",x,x2=x2,out_shape=S,iters=iters,sub=0,act=1); print("binary iters",iters,round(ms*1e3,1),"us",flush=True)
binary(); binary(); unary(0); binary(); binary(1); binary(2); binary(50); unary(1); binary(); binary()
