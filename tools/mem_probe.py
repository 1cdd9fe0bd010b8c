import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from relagn_b200.model import relagn
torch.cuda.init()
free0, tot = torch.cuda.mem_get_info()
m = relagn()
s = m.get_totSED(rel=True)
torch.cuda.synchronize()
free1, _ = torch.cuda.mem_get_info()
print("single object: device memory taken %.1f MB (incl. CUDA context growth, tables)" % ((free0 - free1) / 1e6))
