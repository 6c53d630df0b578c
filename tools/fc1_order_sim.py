#!/usr/bin/env python
"""CPU replay (numpy, no GPU) of two geometric quantities of the shared-footprint path on the exact cfg2 bench batch:

  * the K-cell-sparse fc1: for a row order (sort key) and tile height, the cells each M tile visits (executed cell-tiles) and the
    union of cells over a band of concurrently running tiles (the weight working set, 8 MB per cell) - DESIGN.md §3a
  * the block-sparse conv3_1: the share of the conv3_1 pixels visited for block shapes of 2x2, 2x1, 1x2 and 1x1 pooled cells

  python tools/fc1_order_sim.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scene_graph_commonsense_b200 import pipeline
S = bench.make_samples(0, with_maps=False)
rects=[]
for s in S:
    b=np.clip(s.bbox.numpy().astype(np.int64),0,32)
    xa,xb=pipeline._cell_interval(b[:,0],b[:,1]); ya,yb=pipeline._cell_interval(b[:,2],b[:,3])
    n=len(b)
    for i in range(n):
        for j in range(n):
            if i==j: continue
            x0,x1=max(xa[i],xa[j]),min(xb[i],xb[j]); y0,y1=max(ya[i],ya[j]),min(yb[i],yb[j])
            if x1>x0 and y1>y0: rects.append((x0,x1-1,y0,y1-1))
            else: rects.append((-1,-1,-1,-1))
R=np.array(rects); print(len(R), (R[:,0]<0).mean())
def mask(r):
    if r[0]<0: return 0
    m=0
    for y in range(r[2],r[3]+1):
        for x in range(r[0],r[1]+1): m|=1<<(y*8+x)
    return m
M=np.array([mask(r) for r in R],dtype=np.uint64)
def pop(x): return bin(int(x)).count('1')
def evaluate(order,name,tile=256,band=9):
    Ms=M[order]; nt=(len(Ms)+tile-1)//tile
    tm=[np.bitwise_or.reduce(Ms[t*tile:(t+1)*tile]) for t in range(nt)]
    ex=sum(pop(x) for x in tm)
    # sliding window of `band` tiles: union cells
    un=[pop(np.bitwise_or.reduce(np.array(tm[t:t+band],dtype=np.uint64))) for t in range(0,nt,band)]
    # W traffic model: per band union cells * 8MB
    print("%-28s tiles %d executed cells %d  band-union mean %.1f max %d  sum(band union) %d -> W %.1f GB" % (name,nt,ex,np.mean(un),max(un),sum(un),sum(un)*8/1024))
valid=R[:,0]>=0
def key_cur(r): return np.where(r[:,0]<0, 4096, ((r[:,2]*8+r[:,3])*8+r[:,0])*8+r[:,1])
evaluate(np.argsort(key_cur(R),kind='stable'),"current (y0,y1,x0,x1)")
def key2(r): return np.where(r[:,0]<0, 1<<20, ((r[:,2]*8+r[:,0])*8+r[:,3])*8+r[:,1])
evaluate(np.argsort(key2(R),kind='stable'),"(y0,x0,y1,x1)")
# centre-based Morton / by first cell then last cell
def key3(r): return np.where(r[:,0]<0, 1<<20, ((r[:,2]*8+r[:,0])*64 + (r[:,3]*8+r[:,1])))
evaluate(np.argsort(key3(R),kind='stable'),"(lo cell, hi cell)")
# boustrophedon on (y0,x0) then size
def key4(r):
    x0=np.where(r[:,2]%2==1,7-r[:,0],r[:,0])
    return np.where(r[:,0]<0,1<<20,((r[:,2]*8+x0)*8+(r[:,3]-r[:,2]))*8+(r[:,1]-r[:,0]))
evaluate(np.argsort(key4(R),kind='stable'),"snake (y0,x0) then h,w")
# by centre (2*cy,2*cx) snake then size
def key5(r):
    cy=r[:,2]+r[:,3]; cx=r[:,0]+r[:,1]
    cxs=np.where(cy%2==1,14-cx,cx)
    return np.where(r[:,0]<0,1<<20,((cy*16+cxs)*8+(r[:,3]-r[:,2]))*8+(r[:,1]-r[:,0]))
evaluate(np.argsort(key5(R),kind='stable'),"snake centre then h,w")
for band in (4,9,18):
    evaluate(np.argsort(key_cur(R),kind='stable'),"current band=%d"%band,band=band)
evaluate(np.argsort(key_cur(R),kind='stable'),"current tile=512",tile=512,band=5)
evaluate(np.argsort(key5(R),kind='stable'),"centre tile=512",tile=512,band=5)
print("---- conv3 cover fractions")
v=R[R[:,0]>=0]; w=v[:,1]-v[:,0]+1; h=v[:,3]-v[:,2]+1; P=len(R)
c=lambda a,b: -(-a//b)
print("1x1 cells", (w*h).sum()/64/P)
print("2x2 cells", (c(w,2)*c(h,2)*4).sum()/64/P)
print("2x1 cells (4px wide x 2px tall)", (c(w,2)*2*h).sum()/64/P)
print("1x2 cells", (w*c(h,2)*2).sum()/64/P)
print("best of 2x1/1x2 per pair", np.minimum(c(w,2)*2*h, w*c(h,2)*2).sum()/64/P)
print("mixed: 2x2 then 2x1/1x2/1x1 remainder (exact cover)", (w*h).sum()/64/P)


def schedule_imbalance(order, n_workers, group_m, tile=256, n_tiles_n=16, epi=0.6):
    """Static round-robin persistent schedule of tc_gemm (tile t -> worker t % n_workers, bands of group_m M tiles x all N tiles, n slow
    inside a band): max / mean of the per-worker sums of (cells visited + `epi` cell-equivalents of epilogue) per tile."""
    Ms = M[order]
    nt = (len(Ms) + tile - 1) // tile
    cost = np.array([pop(np.bitwise_or.reduce(Ms[t * tile:(t + 1) * tile])) for t in range(nt)], dtype=np.float64)
    load = np.zeros(n_workers)
    t = 0
    for band in range(0, nt, group_m):
        gm = min(group_m, nt - band)
        for i in range(gm * n_tiles_n):
            load[t % n_workers] += cost[band + i % gm] + epi
            t += 1
    return load.max() / load.mean(), load.max(), load.mean()


print("---- fc1 static-schedule imbalance (max/mean of per-worker work)")
o = np.argsort(key_cur(R), kind='stable')
for workers, gm in ((148, 9), (74, 4), (74, 9), (74, 1)):
    print("workers %d group_m %d: max/mean %.3f (max %.1f mean %.1f cell-steps)" % ((workers, gm) + schedule_imbalance(o, workers, gm)))


def schedule_imbalance_lpt(order, n_workers, group_m, tile=256, n_tiles_n=16, epi=0.6):
    """The same static schedule with the M tiles visited in DESCENDING cost order (longest first)."""
    Ms = M[order]
    nt = (len(Ms) + tile - 1) // tile
    cost = np.sort(np.array([pop(np.bitwise_or.reduce(Ms[t * tile:(t + 1) * tile])) for t in range(nt)], dtype=np.float64))[::-1]
    load = np.zeros(n_workers)
    t = 0
    for band in range(0, nt, group_m):
        gm = min(group_m, nt - band)
        for i in range(gm * n_tiles_n):
            load[t % n_workers] += cost[band + i % gm] + epi
            t += 1
    return load.max() / load.mean(), load.max(), load.mean()


for workers, gm in ((74, 4), (74, 9), (74, 1)):
    print("longest-first: workers %d group_m %d: max/mean %.3f (max %.1f mean %.1f)" % ((workers, gm) + schedule_imbalance_lpt(o, workers, gm)))
