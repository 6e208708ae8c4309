import sys, numpy as np
sys.path.insert(0,'.')
import rtvqa_b200
from rtvqa_b200 import _native as N
from oracle import np_oracle as NO, c_oracle as CO
clip=np.load('tests/golden/small_clip.npz')['clip']
ctx=N.Context(0)
f=clip[0]
win,sc,cnt=ctx.debug_orb(f)
g=NO.bgr2gray(NO.resize_linear_u8(f,64,64)).astype(int)
print("win equal", np.array_equal(win,g[27:37,27:37])); print(win-g[27:37,27:37]); print(sc, cnt, CO.orb_count_64(g.astype(np.uint8)))
