import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import streamly_lz4_b200 as lz
from streamly_lz4_b200 import datagen
total=1<<30; BLOCK=640000
host=datagen.make("mixed",2,total)
offs=np.arange(0,total,BLOCK,dtype=np.int64); lens=np.minimum(BLOCK,total-offs).astype(np.int32)
ctx=lz.Context(0)
src=ctx.pinned("s",total); src[:total]=host
dst=ctx.pinned("d",int((lens.astype(np.int64)+lens//255+24).sum()))
for i in range(4):
    if i==3: os.environ["B200LZ4_DEBUG"]="1"
    t=time.perf_counter(); rc,doff,ol=ctx.compress_batch(src[:total],offs,lens,400,8,dst); dt=time.perf_counter()-t
    print("wall ms",dt*1e3, ctx.timing())
