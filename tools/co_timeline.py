"""Phase timeline (clock64 of CTA 0, second pair) of the tcgen05 co-attention kernels."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, gcnbmp
H, O, mb, N = 128, 128, 2048, 64
rng = np.random.default_rng(0)
a1 = torch.tensor(rng.standard_normal((mb, N, H)) * 0.5, dtype=torch.float32, device="cuda")
a2 = torch.tensor(rng.standard_normal((mb, N, H)) * 0.5, dtype=torch.float32, device="cuda")
link = gcnbmp.NieFineCoattention(H, O, 8, activation=gcnbmp.functions.tanh)
link.mode = gcnbmp.MODE_BF16
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
lib = gcnbmp._capi.lib
lib.bmp_debug_set_buffer_ctc.argtypes = [C.c_void_p]
names = ["X load", "G1 wait", "G1 epi(+S)", "G2 wait", "G2 epi", "stats+L", "H1/H2", "scores+softmax", "pool", "compact|dp",
         "dattn+softmax bwd", "dpre", "u+dlt", "dC", "rsum/cs", "dV+panels", "R wait", "R epi", "A2 wait", "A2 epi+A1 wait", "A1 epi"]
with torch.no_grad():
    link(a1, None, a2, None)
    lib.bmp_debug_set_buffer_ctc(C.c_void_p(dbg.data_ptr()))
    link(a1, None, a2, None)
torch.cuda.synchronize()
d = dbg.cpu().numpy()
print("forward :", "  ".join("%s %d" % (names[i], d[i + 1] - d[i]) for i in range(10)), " total", d[10] - d[0])
dbg.zero_()
t1, t2 = a1.clone().requires_grad_(), a2.clone().requires_grad_()
p1, p2 = link(t1, None, t2, None)
(p1.sum() + p2.sum()).backward()
torch.cuda.synchronize()
d = dbg.cpu().numpy()
print("backward:", "  ".join("%s %d" % (names[i], d[i + 1] - d[i]) for i in range(21)), " total", d[21] - d[0])
