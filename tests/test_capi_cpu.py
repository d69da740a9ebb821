"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/gcnbmp.h declares, and the ctypes structures mirror the C structs byte for byte
(sizes/offsets checked against a gcc-compiled probe).  No compute calls here."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gcnbmp.h")


def test_library_loads_and_exports_every_declared_symbol():
    from gcnbmp import _capi
    text = open(HEADER).read()
    declared = sorted(set(re.findall(r"\b(bmp_[a-z0-9_]+)\s*\(", text)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(_capi.lib, name), "libgcnbmp.so does not export %s" % name
    assert sorted(_capi.EXPORTS) == declared
    assert _capi.lib.bmp_version() >= 100


def test_ctypes_structs_mirror_the_header():
    from gcnbmp import _capi
    structs = {"bmp_gru_t": _capi.GRU, "bmp_ggnn_fwd_t": _capi.GgnnFwd, "bmp_ggnn_bwd_t": _capi.GgnnBwd,
               "bmp_relgcn_fwd_t": _capi.RelgcnFwd, "bmp_relgcn_bwd_t": _capi.RelgcnBwd,
               "bmp_readout_fwd_t": _capi.ReadoutFwd, "bmp_readout_bwd_t": _capi.ReadoutBwd,
               "bmp_coattn_fwd_t": _capi.CoattnFwd, "bmp_coattn_bwd_t": _capi.CoattnBwd, "bmp_bimpm_t": _capi.Bimpm,
               "bmp_pair_t": _capi.Pair}
    probes = []
    for cname, cls in structs.items():
        probes.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            probes.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "gcnbmp.h"\nint main(void){%s return 0;}\n' % "\n".join(probes)
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "probe")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    got = dict(line.split() for line in out.strip().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_missing_library_fails_loudly(tmp_path):
    """The product path has no CPU fallback: without libgcnbmp.so the import raises."""
    pkg = tmp_path / "gcnbmp"
    src = os.path.join(ROOT, "gcn-bmp_b200", "gcnbmp")
    pkg.mkdir()
    for f in os.listdir(src):
        if f.endswith(".py"):
            (pkg / f).write_text(open(os.path.join(src, f)).read())
    code = "import sys; sys.path.insert(0, %r); import gcnbmp" % str(tmp_path)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_ops_refuse_cpu_tensors():
    import torch
    import gcnbmp
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gcnbmp.HolE(1, ()).circular_correlation(torch.zeros(2, 8), torch.zeros(2, 8)) if False else \
            gcnbmp.functional.HoleCorr.apply(torch.zeros(2, 8), torch.zeros(2, 8))


def test_workspace_size_queries_are_host_only_and_consistent():
    """The BF16-mode sizing functions run without a GPU: non-zero exactly for the shapes the tcgen05 kernels cover."""
    from gcnbmp import _capi
    lib = _capi.lib
    for H in (64, 128):
        assert lib.bmp_ggnn_tc_workspace_bytes(H, 6) > 11 * (H // 64) * H * 128 * 6
        assert lib.bmp_relgcn_tc_workspace_bytes(H, 4) > 5 * (H // 64) * H * 128 * 4
        assert lib.bmp_coattn_tc_workspace_bytes(H) > (H // 64) * (H + 32) * 128
        assert lib.bmp_readout_tc_workspace_bytes(H, H) > 0
        # bf16 panel stash: 448 KB (H=128) / 224 KB (H=64) per (step, 2-molecule tile)
        per_unit = (3 + 3 + 4) * (H // 64) * 16384 + 4 * 128 * H * 2
        assert lib.bmp_ggnn_stash2_bytes(2048, H, 6) >= per_unit * 1024 * 6
        assert lib.bmp_ggnn_stash2_bytes(2047, H, 6) == lib.bmp_ggnn_stash2_bytes(2048, H, 6)      # odd batch: padded tile
    # hidden 256: forward-only encoder / readout kernels (weight images + a 64 KB adjacency image per CTA), no stash
    assert lib.bmp_ggnn_tc_workspace_bytes(256, 8) > 8 * 176 * 8192 + 148 * 65536
    assert lib.bmp_readout_tc_workspace_bytes(256, 256) >= 32 * 256 * 128
    assert lib.bmp_readout_tc_workspace_bytes(256, 128) == 0
    for H in (16, 32, 96, 256):
        assert lib.bmp_ggnn_tc_workspace_bytes(H, 6) == 0 or H == 256
        assert lib.bmp_relgcn_tc_workspace_bytes(H, 4) == 0
        assert lib.bmp_coattn_tc_workspace_bytes(H) == 0
        assert lib.bmp_ggnn_stash2_bytes(64, H, 6) == 0


def test_fp32_tensor_core_workspace_query_is_host_only():
    """csrc/ggnn_x3.cu: the BMP_MODE_F32 encoder runs on tcgen05 only when the caller hands it this workspace; the size query
    runs without a GPU and is 0 for the shapes that stay on the FFMA kernels."""
    lib = __import__("gcnbmp")._capi.lib
    q = lib.bmp_ggnn_x3_workspace_bytes
    for H in (64, 128, 256):
        n = q(4144, 64, H, 4, 6, 0)
        rows = 4144 * 64
        assert n > rows * 4 * H * 4 + rows * 64 * 4                     # per-step temporaries (A_e h for 4 bond types, degrees)
        assert q(4144, 64, H, 4, 6, 1) >= n + rows * 7 * H * 4          # inference: + a one-step stash
        assert q(2 * 4144, 64, H, 4, 6, 0) > n
    assert q(4144, 64, 32, 4, 6, 0) == 0 and q(4144, 64, 96, 4, 6, 0) == 0 and q(4144, 64, 128, 3, 6, 0) == 0
    assert q(1, 64, 128, 4, 6, 0) == 0                                  # fewer rows than one 128-row tile
    assert q(4144, 64, 128, 4, 17, 0) == 0
    # molecules with more than 64 atoms: covered up to 256 atoms while an N x hidden fp32 tile fits the adjacency kernels
    assert q(100, 128, 128, 4, 3, 0) > 0 and q(100, 256, 128, 4, 3, 0) > 0 and q(100, 100, 256, 4, 3, 0) > 0
    assert q(100, 257, 128, 4, 3, 0) == 0 and q(100, 256, 256, 4, 3, 0) == 0


def test_fp32_readout_tensor_core_workspace_query_is_host_only():
    q = __import__("gcnbmp")._capi.lib.bmp_readout_x3_workspace_bytes
    n = q(4144, 64, 128, 128, 2, 1)                                      # R2 with h0: i over [h | h0], j over h
    assert n >= 2 * 4144 * 64 * 128 * 4 + (2 + 1) * 2 * 32768              # two pre-activation arrays + 6 packed hi/lo k-tiles
    assert q(4144, 64, 128, 128, 1, 1) > n                                 # R1: j sees [h | h0] too
    assert q(4144, 64, 128, 128, 3, 1) == 0 and q(4144, 64, 96, 128, 1, 1) == 0 and q(4144, 64, 128, 40, 1, 1) == 0
    assert q(1, 64, 128, 128, 1, 1) == 0                                   # fewer rows than one tile


def test_pair_step_workspace_query_is_host_only():
    """bmp_pair_workspace_bytes (csrc/pair.cu) sizes the ONE caller-owned buffer of bmp_pair_forward_backward without a GPU."""
    q = __import__("gcnbmp")._capi.lib.bmp_pair_workspace_bytes
    mb, n1, n2, H, O, hd, K, T = 4144, 64, 64, 128, 128, 8, 86, 6
    f32, bf16 = q(mb, n1, n2, H, O, hd, K, T, 0), q(mb, n1, n2, H, O, hd, K, T, 1)
    rows = mb * 64
    stash = (T + 1) + T + 3 * T + T + 4 * T + (T + 1)                   # Hs, Ms, Gs, RSs, Ps, dHs in units of rows x H floats
    assert f32 > 2 * stash * rows * H * 4                               # both drugs' fp32 stashes (+ the tensor-core workspace)
    assert 0 < bf16 < f32                                               # the bf16 panel stash is the smaller tape
    assert q(mb, n1, n2, 96, O, hd, K, T, 1) == 0                       # no bf16 panel stash at hidden 96
    assert q(mb, n1, n2, 96, O, hd, K, T, 0) > 0                        # fp32 mode: FFMA encoders cover it
    assert q(0, n1, n2, H, O, hd, K, T, 0) == 0 and q(mb, n1, n2, H, O, hd, K, 17, 0) == 0


def test_chainer_adapter_is_import_guarded():
    """the Chainer / CuPy adapter imports without either package and fails loudly, not silently, when used."""
    from gcnbmp import chainer_adapter as B
    if B.AVAILABLE:          # a Chainer-equipped process: nothing to guard
        return
    import pytest
    with pytest.raises(ImportError):
        B.ggnn_encode(None, None, None)
    with pytest.raises(ImportError):
        B.GGNNEncode(None, None, [], 0, 0)
