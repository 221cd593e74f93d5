import sys, torch; sys.path.insert(0, "."); sys.path.insert(0, "tools")
import flow_supervisor_b200 as fsb
from flow_supervisor_b200 import ops
from bench_rows import timed
B, D, H, W = 8, 256, 55, 128
g = torch.Generator().manual_seed(0)
f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda(); f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
c = (fsb.coords_grid(B, H, W) + 5.0 * torch.randn(B, 2, H, W, generator=g)).cuda()
conv = torch.nn.Conv2d(324, 256, 1).cuda()
packed = ops.convc1_prepare(conv.weight.detach(), conv.bias.detach())
for vol in ("f32", "bf16"):
    fsb.CorrBlock.volume = vol
    try:
        blk = fsb.CorrBlock(f1, f2)
        with torch.no_grad():
            print(vol, "lookup us", 1e3 * timed(lambda: blk(c), 5, inner=12), "fused us", 1e3 * timed(lambda: blk.lookup_convc1(c, packed), 5, inner=12))
    except Exception as e:
        print(vol, "error", str(e)[:200])
