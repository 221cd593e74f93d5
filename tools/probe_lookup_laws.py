"""Lookup forward time by coordinate law (config 2 geometry): random flow (the bench), the lattice of GRU iteration 0
(integer coordinates: floor flips -> per-tap path), half-pixel offsets, large out-of-bounds flow."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_supervisor_b200 as fsb
B, D, H, W = 8, 256, 55, 128
g = torch.Generator().manual_seed(0)
f1 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda(); f2 = (1.57 * torch.randn(B, D, H, W, generator=g)).cuda()
grid = fsb.coords_grid(B, H, W)
laws = {"random N(0,5^2)": grid + 5.0 * torch.randn(B, 2, H, W, generator=g), "lattice (iteration 0)": grid.clone(),
        "lattice + integer flow": grid + torch.randint(-4, 5, (B, 2, H, W), generator=g).float(),
        "half pixel": grid + 0.5, "smooth small flow": grid + 0.3 * torch.randn(B, 2, H, W, generator=g),
        "out of bounds N(0,40^2)": grid + 40.0 * torch.randn(B, 2, H, W, generator=g)}
blk = fsb.CorrBlock(f1, f2)
for name, c in laws.items():
    c = c.cuda()
    for _ in range(3): blk(c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): blk(c)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"law": name, "lookup_us": e0.elapsed_time(e1) / 20 * 1e3}))
