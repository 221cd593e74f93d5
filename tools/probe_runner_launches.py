"""ncu launch-list driver: ONE replay of the RaftRunner graph (reference RAFT, 8 pairs, 436x1024, 12 iterations)."""
import argparse, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flow_supervisor_b200 as fsb
from baseline import install_ref
sys.path.insert(0, install_ref.path())
from core.raft import RAFT
torch.manual_seed(1234)
model = RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False)).eval().cuda()
g = torch.Generator().manual_seed(0)
im1 = (torch.rand(8, 3, 440, 1024, generator=g) * 255).cuda(); im2 = (torch.rand(8, 3, 440, 1024, generator=g) * 255).cuda()
run = fsb.RaftRunner(model, iters=12, graph=os.environ.get("RUNNER_GRAPH", "0") == "1")
with torch.no_grad():
    for _ in range(2): run(im1, im2)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run(im1, im2)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
