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
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
lib = gcnbmp._capi.lib
lib.bmp_debug_set_buffer_co.argtypes = [C.c_void_p]
with torch.no_grad():
    link(a1, None, a2, None)
    lib.bmp_debug_set_buffer_co(C.c_void_p(dbg.data_ptr()))
    link(a1, None, a2, None)
torch.cuda.synchronize()
d = dbg.cpu().numpy()
names = ["load", "v+Q gemm", "C gemm", "stats+L", "lt proj", "H1/H2", "scores", "attn softmax", "pool"]
print("  ".join("%s %d" % (names[i], d[i + 1] - d[i]) for i in range(8)), " total", d[8] - d[0])
