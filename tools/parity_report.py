"""Print the SegModel parity numbers (engine on cuda:0 vs the fp32 CPU oracle) for the three synthetic plans.
    python tools/parity_report.py [--bwd] [--c1]          (REHR_FWD_DTYPE=bf16 selects the all-bf16 operand mode)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import parity  # noqa: E402
from rehrseg_b200 import functional as Fn  # noqa: E402

bwd = "--bwd" in sys.argv
cases = [("tiny", (16, 32, 32), 2), ("3d_fullres", (64, 64, 64), 1), ("anisotropic", (8, 64, 64), 1), ("anisotropic", (16, 128, 128), 1)]
if "--c1" in sys.argv:
    cases = [("3d_fullres", (128, 128, 128), 2)]
print("FWD_FP16 =", Fn.FWD_FP16)
for plan, patch, b in cases:
    res = parity.segmodel_parity(patch=patch, batch=b, plan=plan, backward=bwd, threads=os.cpu_count())
    pp = res.pop("per_param_grad_rel_l2", None)
    if pp:
        top = sorted(pp.items(), key=lambda kv: -kv[1])[:6]
        res["worst_params"] = {k: round(v, 4) for k, v in top}
    print(plan, patch, b, json.dumps({k: (round(v, 6) if isinstance(v, float) else v) for k, v in res.items()}))
    torch.cuda.empty_cache()
