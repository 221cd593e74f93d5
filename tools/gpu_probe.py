"""One-off probes of facts the survey could not check without a GPU (SURVEY.md hard
part 1): how ATen rounds tensor / python_int on CUDA, and whether grid_sample routes to
cuDNN.  Output goes to stdout (captured into gpurun_out/)."""
import numpy as np
import torch

torch.manual_seed(0)
x = (torch.arange(0, 4000, dtype=torch.float32) * 0.5).cuda()
for den in (61, 95, 127, 155, 54, 46, 239, 135, 30, 15, 7):
    t = 2 * x
    q = (t / den).cpu().numpy()
    true_div = (t.cpu().numpy() / np.float32(den)).astype(np.float32)
    recip = (t.cpu().numpy() * (np.float32(1) / np.float32(den))).astype(np.float32)
    print(f"den={den:4d}  ==true_div: {int((q == true_div).sum()):5d}/{q.size}  ==recip_mul: {int((q == recip).sum()):5d}/{q.size}")

# does grid_sample go through cuDNN? compare kernels via the profiler
import torch.nn.functional as F
inp = torch.randn(64, 1, 27, 64, device="cuda")
grid = torch.rand(64, 9, 9, 2, device="cuda") * 2 - 1
for flag in (True, False):
    torch.backends.cudnn.enabled = flag
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        F.grid_sample(inp, grid, align_corners=True)
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    print("cudnn.enabled =", flag, "->", names)
torch.backends.cudnn.enabled = True
print(torch.cuda.get_device_name(0), torch.version.cuda, torch.backends.cudnn.version())
