"""Development tool: kernel timeline (start, duration, stream) of one graph-replayed C1 step from torch.profiler / CUPTI,
written as CSV (gpurun_out/timeline.csv by default).  `tools/timeline_report.py` turns it into an exposed-time table."""
import sys, os, argparse, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from rehrseg_b200 import seg_model as sm
from rehrseg_b200.graphs import GraphedTrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/timeline.csv")
a = ap.parse_args()
torch.manual_seed(0)
m = sm.plainconv_unet_3d_fullres().cuda()
x = torch.randn(2, 1, 128, 128, 128, device="cuda")
g = torch.randn(2, 2, 128, 128, 128, device="cuda")
gs = GraphedTrainStep(m, lambda out, g: torch.dot(out.float().reshape(-1), g.reshape(-1)) / out.numel(), (x, g))
for _ in range(5):
    gs.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        gs.replay()
    torch.cuda.synchronize()
tmp = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tmp)
ev = json.load(open(tmp))["traceEvents"]
ks = [e for e in ev if e.get("cat") == "kernel"]
ks.sort(key=lambda e: e["ts"])
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
with open(a.out, "w") as f:
    f.write("ts_us,dur_us,stream,name\n")
    for e in ks:
        f.write(f"{e['ts']:.3f},{e['dur']:.3f},{e['args'].get('stream', -1)},\"{e['name'][:110]}\"\n")
print("kernels:", len(ks), "->", a.out)
