"""128-byte-line model of the lookup forward's DRAM reads at config 2 (B=8, 55x128 tokens, bench coordinates): distinct
lines touched per query and level by the 10x10 footprints, for the shipped layout (2x8 patches, row-pair major: a line =
2 rows x 16 columns) and for line-shaped alternatives (4x8, 8x4; bf16: 2x32 shipped, 4x16).  The shipped layout comes out at
165.7 MB per launch -- ncu measures 166.3 MB.  See DESIGN.md section 2 and profiles/r02h_line_traffic_model.txt."""
import numpy as np, torch
H,W,B,L,R=55,128,8,4,4
g=torch.Generator().manual_seed(0)
_=torch.randn(B,256,H,W,generator=g); _=torch.randn(B,256,H,W,generator=g)
ys,xs=torch.meshgrid(torch.arange(H),torch.arange(W),indexing="ij")
grid=torch.stack([xs,ys],0).float()[None]
c=(grid+5.0*torch.randn(B,2,H,W,generator=g)).numpy()
Q=B*H*W
def clipcount(lo,hi,mx):  # number of integer cells in [lo,hi] ∩ [0,mx-1]
    return np.clip(np.minimum(hi,mx-1)-np.maximum(lo,0)+1,0,None)
tot_old=tot_new=tot_new8x4=tot_bf_old=tot_bf_new=0
for l in range(L):
    Hl,Wl=H>>l,W>>l; Wp=(Wl+7)//8*8
    cx=np.floor(c[:,0]/2**l).astype(int); cy=np.floor(c[:,1]/2**l).astype(int)
    xl=cx-R; xh=cx+R+1; yl=cy-R; yh=cy+R+1     # inclusive element ranges (10 wide)
    # old: lines = 2 rows x 16 cols (when Wp%16==0)
    Hp2=(Hl+1)//2*2
    old=clipcount(yl>>1,yh>>1,Hp2//2)*clipcount(xl>>4,xh>>4,(Wp+15)//16)
    Hp4=(Hl+3)//4*4
    new=clipcount(yl>>2,yh>>2,Hp4//4)*clipcount(xl>>3,xh>>3,Wp//8)
    n84=clipcount(yl>>3,yh>>3,(Hl+7)//8)*clipcount(xl>>2,xh>>2,(Wl+3)//4)
    # bf16: old line = 2 rows x 32 cols ; new bf16 line 4 x 16 (same patch order, 2B elements)
    bo=clipcount(yl>>1,yh>>1,Hp2//2)*clipcount(xl>>5,xh>>5,(Wp+31)//32)
    bn=clipcount(yl>>2,yh>>2,Hp4//4)*clipcount(xl>>4,xh>>4,(Wp+15)//16)
    print(l,"old lines",old.mean(),"new 4x8",new.mean(),"8x4",n84.mean(),"bf16 old",bo.mean(),"bf16 new",bn.mean())
    tot_old+=old.mean();tot_new+=new.mean();tot_new8x4+=n84.mean();tot_bf_old+=bo.mean();tot_bf_new+=bn.mean()
for n,t in (("old",tot_old),("new 4x8",tot_new),("8x4",tot_new8x4),("bf16 old",tot_bf_old),("bf16 new(4x16)",tot_bf_new)):
    print(n,"MB/launch",t*128*Q/1e6)
