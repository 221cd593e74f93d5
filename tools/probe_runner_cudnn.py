"""RaftRunner forward time with torch.backends.cudnn.benchmark off / on and under bf16 autocast (model option
mixed_precision): how much of the model-level time is cuDNN's algorithm choice (8 pairs 436x1024, 12 iterations)."""
import argparse, json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import flow_supervisor_b200 as fsb
from baseline import install_ref
sys.path.insert(0, install_ref.path())
from core.raft import RAFT
from bench_rows import timed
g = torch.Generator().manual_seed(0)
im1 = (torch.rand(8, 3, 440, 1024, generator=g) * 255).cuda(); im2 = (torch.rand(8, 3, 440, 1024, generator=g) * 255).cuda()
for name, bench, mp in (("default", False, False), ("cudnn.benchmark", True, False), ("mixed_precision", False, True), ("mixed_precision + cudnn.benchmark", True, True)):
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(1234)
    model = RAFT(argparse.Namespace(small=False, mixed_precision=mp, alternate_corr=False)).eval().cuda()
    run = fsb.RaftRunner(model, iters=12, graph=True)
    with torch.no_grad():
        ms = timed(lambda: run(im1, im2), 4, warm=2)
    print(json.dumps({"variant": name, "ms": ms, "pairs_per_s": 8 / ms * 1e3}))
    del run, model; torch.cuda.empty_cache()
