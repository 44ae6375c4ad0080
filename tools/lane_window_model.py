"""CPU model (not a GPU measurement) of the WINDOW traversal of csrc/blend_backward.cu on the bench workload: box-culled
survivors of a warp's 8x4 rectangle are processed KW at a time; after the lock-step evaluate phase every lane walks only
the slots that blend ITS pixel, so the chain phase of a window costs max-over-lanes(contributing slots) iterations
instead of KW.  Prints, per warp: box hits (= iterations of a warp-lock-stepped walk), contributing (lane, slot) pairs, and
chain iterations for KW in {16, 32, 64, 128} and without windowing ("free").  Built on the C oracle's geometry + binning
and tools/lane_packing_model.py's cull boxes.   Run: python tools/lane_window_model.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import scenes
from lane_packing_model import cull_half_extents
sc = scenes.config3(P=1_000_000)
o = sc.oracle(); geo, b = o.geometry(), o.binning()
xy, co, radii = geo["means2D"], geo["conic_opacity"], geo["radii"]
hx, hy = cull_half_extents(co[:, :3], co[:, 3], radii)
ranges, plist = b["ranges"], b["point_list"]
gx = (sc.w + 15) // 16
rng = np.random.default_rng(0)
nonempty = np.nonzero(ranges[:, 1] > ranges[:, 0])[0]
sample = rng.choice(nonempty, size=60, replace=False)
px = np.arange(16, dtype=np.float32)
tot = dict(hits_b=0, hits_f=0, act=0)
for Wn in (16, 32, 64, 128): tot[f'it_b{Wn}']=0; tot[f'it_f{Wn}']=0; tot[f'win_b{Wn}']=0
tot['it_b_free']=0
for t in sample:
    lo, hi = int(ranges[t,0]), int(ranges[t,1]); ids = plist[lo:hi]; n = ids.size
    tx, ty = int(t % gx), int(t // gx)
    X = (tx*16+px)[None,None,:]; Y = (ty*16+px)[None,:,None]
    gxy, gco = xy[ids], co[ids]
    dx = gxy[:,0][:,None,None]-X; dy = gxy[:,1][:,None,None]-Y
    power = -0.5*(gco[:,0][:,None,None]*dx*dx + gco[:,2][:,None,None]*dy*dy) - gco[:,1][:,None,None]*dx*dy
    alpha = np.minimum(0.99, gco[:,3][:,None,None]*np.exp(np.minimum(power,0.0)))
    contrib = (power<=0)&(alpha>=1/255.); valid=(Y<sc.h)&(X<sc.w); contrib&=valid
    one_m = np.where(contrib,1-alpha,1.0).astype(np.float64); T_incl=np.cumprod(one_m,axis=0)
    stop = contrib&(T_incl<1e-4); stopped=np.cumsum(stop,axis=0)>0; live=contrib&~stopped
    evalf = contrib & ~(np.cumsum(stop,axis=0)-stop>0)   # forward evaluates up to and including the stopping pair
    idx1=np.arange(1,n+1)[:,None,None]; n_contrib=np.max(np.where(live,idx1,0),axis=0)
    done_at=np.where(stopped.any(axis=0),np.argmax(stopped,axis=0)+1,n)
    hxs,hys=hx[ids],hy[ids]
    for w in range(8):
        x0,y0=tx*16+(w&1)*8, ty*16+(w>>1)*4
        sl=(slice(None),slice((w>>1)*4,(w>>1)*4+4),slice((w&1)*8,(w&1)*8+8))
        fwd_n=int(done_at[sl[1],sl[2]].max()); bwd_n=int(n_contrib[sl[1],sl[2]].max())
        box=(gxy[:,0]+hxs>=x0)&(gxy[:,0]-hxs<=x0+7)&(gxy[:,1]+hys>=y0)&(gxy[:,1]-hys<=y0+3)
        actb=live[sl].reshape(n,32); actf=evalf[sl].reshape(n,32)
        sb=np.nonzero(box[:bwd_n])[0][::-1]; sf=np.nonzero(box[:fwd_n])[0]
        tot['hits_b']+=sb.size; tot['hits_f']+=sf.size; tot['act']+=int(actb[sb].sum())
        tot['it_b_free']+=int(actb[sb].sum(axis=0).max()) if sb.size else 0
        for Wn in (16,32,64,128):
            for c0 in range(0,sb.size,Wn):
                tot[f'it_b{Wn}']+=int(actb[sb[c0:c0+Wn]].sum(axis=0).max()); tot[f'win_b{Wn}']+=1
            for c0 in range(0,sf.size,Wn):
                tot[f'it_f{Wn}']+=int(actf[sf[c0:c0+Wn]].sum(axis=0).max())
nw=sample.size*8
print({k: round(v/nw,1) for k,v in tot.items()})
