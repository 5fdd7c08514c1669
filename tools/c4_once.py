"""BASELINE config 4 alone, exactly as bench.py's `configs.c4_joint` entry measures it (bench_extras.c4_joint)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, bench_extras
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
keys = (1,) if len(sys.argv) > 1 and sys.argv[1] == "keys" else None
print(json.dumps(bench_extras.c4_joint(dev, bench.measured_peaks(), teacher_keys=keys)))
