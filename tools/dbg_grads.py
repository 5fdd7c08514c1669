import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import parity
from rehrseg_b200 import functional as Fn
for march in (True, False):
    Fn.USE_MARCH = march
    res = parity.segmodel_parity(patch=(16, 32, 32), batch=2, plan="tiny", backward=True)
    print("march", march, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()})
