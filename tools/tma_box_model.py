"""Host-side model of the lookup forward's TMA traffic at config 2 (B=8, 55x128 tokens, bench coordinates): 64-byte patches
requested per query and level by the {5|6} x {2|3} boxes ("full box"), by boxes clipped in y only and by boxes clipped to the map
on all sides ("both").  The numbers match ncu's l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld to the byte: 168.45 MB unclipped,
117.64 MB clipped (profiles/r02a_tma_box_traffic.txt)."""
import numpy as np, torch
H,W,B,L,R=55,128,8,4,4
g=torch.Generator().manual_seed(0)
_=torch.randn(B,256,H,W,generator=g); _=torch.randn(B,256,H,W,generator=g)
ys,xs=torch.meshgrid(torch.arange(H),torch.arange(W),indexing="ij")
grid=torch.stack([xs,ys],0).float()[None]
c=(grid+5.0*torch.randn(B,2,H,W,generator=g)).numpy()
tot_box=0; tot_need=0; tot_el=0
for l in range(L):
    Hl,Wl=H>>l,W>>l; Wp=(Wl+7)//8*8; Hp=(Hl+1)//2*2
    cx=np.floor(c[:,0]/2**l).astype(int); cy=np.floor(c[:,1]/2**l).astype(int)
    xl=cx-R; xh=cx+R; yl=cy-R; yh=cy+R
    rp0=yl>>1; pc0=xl>>3
    n_rp=((yh+1)>>1)-rp0+1; n_pc=((xh+1)>>3)-pc0+1
    brp=np.where(n_rp>5,6,5); bpc=np.where(n_pc>2,3,2)
    # clipped to map: row pairs [0,Hp/2), patches [0,Wp/8)
    def clip(lo,n,mx): 
        a=np.clip(lo,0,mx); b=np.clip(lo+n,0,mx); return b-a
    box=clip(rp0,brp,Hp//2)*clip(pc0,bpc,Wp//8)
    need=clip(rp0,n_rp,Hp//2)*clip(pc0,n_pc,Wp//8)
    # valid-only patches (exclude pad-only)  
    el=clip(yl,10,Hl)*clip(xl,10,Wl)
    print(l,Hl,Wl,"box patches/query",box.mean(),"needed",need.mean(),"elems",el.mean())
    tot_box+=box.mean()*64; tot_need+=need.mean()*64; tot_el+=el.mean()*4
Q=B*H*W
print("per query bytes: box",tot_box,"need",tot_need,"alg",tot_el, " MB/launch:",tot_box*Q/1e6,tot_need*Q/1e6,tot_el*Q/1e6)
print("---- variants")
tb=0;ty=0;tx=0;tn=0
for l in range(L):
    Hl,Wl=H>>l,W>>l; Wp=(Wl+7)//8*8; Hp=(Hl+1)//2*2
    cx=np.floor(c[:,0]/2**l).astype(int); cy=np.floor(c[:,1]/2**l).astype(int)
    xl=cx-R; xh=cx+R; yl=cy-R; yh=cy+R
    rp0=yl>>1; pc0=xl>>3
    n_rp=((yh+1)>>1)-rp0+1; n_pc=((xh+1)>>3)-pc0+1
    brp=np.where(n_rp>5,6,5); bpc=np.where(n_pc>2,3,2)
    def clip(lo,n,mx):
        a=np.clip(lo,0,mx); b=np.clip(lo+n,0,mx); return b-a
    full=(brp*bpc).mean(); yclip=(clip(rp0,n_rp,Hp//2)*bpc).mean(); both=(clip(rp0,n_rp,Hp//2)*clip(pc0,n_pc,Wp//8)).mean()
    exact=(n_rp*n_pc).mean()
    print(l,"full box",full,"exact unclipped",exact,"y-clip",yclip,"both",both)
    tb+=full;ty+=yclip;tn+=both
print("MB/launch full",tb*64*Q/1e6,"yclip",ty*64*Q/1e6,"both",tn*64*Q/1e6)
