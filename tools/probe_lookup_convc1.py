"""ncu driver: a few launches of the fused lookup + convc1 kernel at config 2 (B=8, 55x128 tokens)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_supervisor_b200 as fsb
from flow_supervisor_b200 import ops
B, D, H, W = 8, 256, 55, 128
g = torch.Generator().manual_seed(0)
f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda(); f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
c = (fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
conv = torch.nn.Conv2d(324, 256, 1).cuda()
packed = ops.convc1_prepare(conv.weight.detach(), conv.bias.detach())
blk = fsb.CorrBlock(f1, f2)
for _ in range(6):
    out = blk.lookup_convc1(c, packed)
torch.cuda.synchronize()
