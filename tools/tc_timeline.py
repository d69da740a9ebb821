"""Phase timeline (SM clock cycles) of CTA 0 of the tcgen05 backward kernel."""
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

H, T, N, mb = 128, 6, 64, 2048
rng = np.random.default_rng(0)
atoms, adj = synthetic.random_molecules(rng, mb, N)
A, X = torch.tensor(adj).cuda(), torch.tensor(atoms).cuda()
net = gcnbmp.GGNNMono(H, H, T)
net.mode = gcnbmp.MODE_BF16
dbg = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
dbgf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
lib = gcnbmp._capi.lib
lib.bmp_debug_set_buffer.argtypes = [C.c_void_p]
lib.bmp_debug_set_buffer_fwd.argtypes = [C.c_void_p]
for i in range(2):
    net.cleargrads()
    if i == 1:
        lib.bmp_debug_set_buffer_fwd(C.c_void_p(dbgf.data_ptr()))
    g = net(X, A)
    lib.bmp_debug_set_buffer_fwd(None)
    if i == 1:
        lib.bmp_debug_set_buffer(C.c_void_p(dbg.data_ptr()))
    (g.sum() + net.get_atom_array().sum()).backward()
torch.cuda.synchronize()
f = dbgf.cpu().numpy().reshape(64, 16)
fn = ["E1 done", "m ready", "E2 done", "r ready", "E3 done", "zh ready", "E4 done"]
print("forward (with stash):")
for it in range(0, 8):
    row = f[it]
    if row[0] == 0:
        break
    print("step %2d: " % it + "  ".join("%s +%d" % (fn[i - 1], row[i] - row[0]) for i in range(1, 8)))
print("backward:")
d = dbg.cpu().numpy().reshape(64, 16)
names = ["A start", "A done", "q ready", "B done", "dx ready", "C done", "P00", "P10", "P01", "P11", "D done", "dh ready", "E done"]
for it in range(0, 14):
    row = d[it]
    if row[0] == 0:
        break
    base = row[0]
    print("step %2d: " % it + "  ".join("%s +%d" % (names[i], row[i] - base) for i in range(1, 13)))
print("step starts relative to the first (forward):", [int(f[i][0] - f[0][0]) for i in range(0, 20) if f[i][0]])
print("step ends   relative to the first (forward):", [int(f[i][7] - f[0][0]) for i in range(0, 20) if f[i][0]])
print("step starts relative to the first (backward):", [int(d[i][0] - d[0][0]) for i in range(0, 20) if d[i][0]])
print("step ends   relative to the first (backward):", [int(d[i][12] - d[0][0]) for i in range(0, 20) if d[i][0]])
print("forward tile prologue (tile 1): ", "  ".join("%s %d" % (n, f[6][8 + i + 1] - f[6][8 + i]) for i, n in enumerate(["embed ld issue", "stage adj", "h0 store", "h operand", "bar", "degrees"])), " from prev E4 done:", f[6][8] - f[5][7], " to step start:", f[6][0] - f[6][14])
print("forward: cycles the MMA lane waited for weight tiles, per step:", [int(f[i][15]) for i in range(0, 12)])
