"""Host check (numpy, no GPU) of the tile algebra of ddh_kernel_reg8 (csrc/ddh.cu): the collocated stiffness of a 2 x 2-element,
n_basis 8 subdomain computed element by element against the four-threads-per-element scheme with mirrored frames, partner tiles
re-indexed 3 - j, negated partner fluxes and x-then-y edge assembly. Prints the maximal difference (rounding level)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import setup_np as S
NB=8; B=S.Basis(NB); x=B.x; wq=B.w
D=B.deriv(x)          # D[j,i] = phi_i'(x_j)  -> Dm[k][i] = D(k,i)
Dm=np.array(D)
print('antisym check', np.abs(Dm[::-1,::-1]+Dm).max())
cx, cz = 1.3, 0.7
g_x=np.outer(wq,wq)*cx   # [l][k]
g_z=np.outer(wq,wq)*cz
N1=15
rng=np.random.default_rng(0)
W=rng.uniform(-1,1,(N1,N1))   # [Y][X]
# reference: per element collocated stiffness, assembled
Z=np.zeros((N1,N1))
for ey in range(2):
    for ex in range(2):
        w=W[ey*7:ey*7+8, ex*7:ex*7+8]      # [l][k]
        Ux=w@Dm.T                           # Ux[l][k] = sum_i D[k][i] w[l][i]
        Uy=Dm@w                             # Uy[l][k] = sum_i D[l][i] w[i][k]
        fx=g_x*Ux; fy=g_z*Uy
        Su=fx@Dm + Dm.T@fy                  # sum_i D[i][k] fx[l][i] + sum_i D[i][l] fy[i][k]
        Z[ey*7:ey*7+8, ex*7:ex*7+8]+=Su
# emulate kernel: 16 threads (tx,ty)
def nk(s,kk): return 7-kk if s else kk
tiles={}
for ty in range(4):
    for tx in range(4):
        ex,ey,sx,sy=tx>>1,ty>>1,tx&1,ty&1
        t=np.zeros((4,4))
        for l in range(4):
            for k in range(4):
                t[l,k]=W[ey*7+nk(sy,l), ex*7+nk(sx,k)]
        tiles[(tx,ty)]=t
def lane(tx,ty): return tx+4*ty
def from_lane(L): return (L&3,(L>>2)&3)
gx4=g_x[:4,:4]; gz4=g_z[:4,:4]
fxs={};fys={}
for (tx,ty),w in tiles.items():
    px=from_lane(lane(tx,ty)^1); py=from_lane(lane(tx,ty)^4)
    wpx=tiles[px][:, ::-1]      # wp[l][j] = partner w[l][3-j]
    wpy=tiles[py][::-1, :]      # wp[j][k] = partner w[3-j][k]
    Ux=np.zeros((4,4));Uy=np.zeros((4,4))
    for l in range(4):
        for k in range(4):
            Ux[l,k]=sum(Dm[k,i]*w[l,i] for i in range(4))+sum(Dm[k,4+i]*wpx[l,i] for i in range(4))
            Uy[l,k]=sum(Dm[l,i]*w[i,k] for i in range(4))+sum(Dm[l,4+i]*wpy[i,k] for i in range(4))
    fxs[(tx,ty)]=gx4*Ux; fys[(tx,ty)]=gz4*Uy
zs={}
for (tx,ty) in tiles:
    px=from_lane(lane(tx,ty)^1); py=from_lane(lane(tx,ty)^4)
    fx=fxs[(tx,ty)]; fy=fys[(tx,ty)]
    fpx=-fxs[px][:, ::-1]; fpy=-fys[py][::-1,:]
    z=np.zeros((4,4))
    for l in range(4):
        for k in range(4):
            z[l,k]=sum(Dm[i,k]*fx[l,i] for i in range(4))+sum(Dm[4+i,k]*fpx[l,i] for i in range(4)) \
                  +sum(Dm[i,l]*fy[i,k] for i in range(4))+sum(Dm[4+i,l]*fpy[i,k] for i in range(4))
    zs[(tx,ty)]=z
# assembly x then y
z2={k:v.copy() for k,v in zs.items()}
for (tx,ty) in zs:
    if tx in (1,2):
        o=from_lane(lane(tx,ty)^3)
        z2[(tx,ty)][:,0]=zs[(tx,ty)][:,0]+zs[o][:,0]
z3={k:v.copy() for k,v in z2.items()}
for (tx,ty) in zs:
    if ty in (1,2):
        o=from_lane(lane(tx,ty)^12)
        z3[(tx,ty)][0,:]=z2[(tx,ty)][0,:]+z2[o][0,:]
err=0
for (tx,ty),z in z3.items():
    ex,ey,sx,sy=tx>>1,ty>>1,tx&1,ty&1
    for l in range(4):
        for k in range(4):
            err=max(err,abs(z[l,k]-Z[ey*7+nk(sy,l), ex*7+nk(sx,k)]))
print('max err', err, 'scale', np.abs(Z).max())
