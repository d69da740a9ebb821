"""clock64() accounting of the hidden-256 encoder (csrc/ggnn_tc256.cu): per step of CTA 0, the MMA-issuer lane's total cycles
and the cycles it spent waiting for weight tiles, AH panels (E1), m panels (E2), r*h panels (E3) and the new state (E4)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gcnbmp
from gcnbmp import synthetic

NP = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lib = gcnbmp._capi.lib
lib.bmp_debug_set_buffer_fwd.argtypes = [C.c_void_p]
rng = np.random.default_rng(1)
atoms, adj = synthetic.random_molecules(rng, NP, 64)
atoms, adj = torch.tensor(atoms).cuda(), torch.tensor(adj).cuda()
net = gcnbmp.GGNN(256, hidden_dim=256, n_layers=8, weight_tying=True)
net.mode = gcnbmp.MODE_BF16
with torch.no_grad():
    net(atoms, adj)
    torch.cuda.synchronize()
    dbg = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
    lib.bmp_debug_set_buffer_fwd(C.c_void_p(dbg.data_ptr()))
    net(atoms, adj)
    torch.cuda.synchronize()
    lib.bmp_debug_set_buffer_fwd(None)
d = dbg.cpu().numpy().reshape(64, 8)
print("step   total  w-wait  AH-wait   x-wait  rs-wait   h-wait")
for i in range(min(24, 64)):
    if d[i, 0] == 0:
        break
    print("%4d %7d %7d %8d %8d %8d %8d" % (i, d[i, 0], d[i, 1], d[i, 2], d[i, 3], d[i, 4], d[i, 5]))
