"""The FlowFormer / Twins token-side kernels (csrc/ops_tokens.cu) through their C-ABI entry points against plain PyTorch fp32 of the
same op — on the B200 (`gpu` marker) and, since every kernel has a thread-per-item functor form, on the CPU build of the same sources
(tests/hostsim; not a product path).  Tolerances are written next to each check."""
import ctypes as C
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import flowformer_oracle as FO
from gimmvfi_b200._lib import Lib, default_lib, view_of

BACKENDS = [pytest.param("hostsim", id="hostsim"), pytest.param("gpu", id="b200", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module", params=BACKENDS)
def be(request):
    if request.param == "gpu":
        return default_lib(), "cuda"
    import harness
    return Lib(harness.build_hostsim()), "cpu"


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream) if t.is_cuda else None


def rnd(*s, seed=0):
    return torch.randn(*s, generator=torch.Generator().manual_seed(seed))


P = lambda t: C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("C_,Hp,Wp,pe", [(128, 0, 0, 0), (256, 3, 5, 0), (128, 0, 0, 4)])
def test_layernorm(be, C_, Hp, Wp, pe):
    lib, dev = be
    n, h, w = 2, 9, 12
    x, g, b = rnd(n, h, w, C_, seed=1), 1 + 0.1 * rnd(C_, seed=2), 0.1 * rnd(C_, seed=3)
    out = torch.full((n, h + Hp, w + Wp, C_), 7.0)
    xd, gd, bd, od = x.to(dev), g.to(dev), b.to(dev), out.to(dev)
    lib.check(lib.dll.gimmvfi_op_layernorm(C.byref(view_of(xd)), P(gd), P(bd), 1e-5, C.byref(view_of(od)), float(pe), C_ if pe else 0, _stream(od)))
    ref = F.layer_norm(x, (C_,), g, b, 1e-5)
    if pe:   # + LinearPositionEmbeddingSine(coords * pe)  (twins.py:500-513)
        co = FO.coords_grid(n, h, w).view(n, 2, -1).permute(0, 2, 1) * pe
        ref = ref + FO.linear_pos_embedding_sine(co, dim=C_).view(n, h, w, C_)
    ref = F.pad(ref, (0, 0, 0, Wp, 0, Hp))
    assert (od.cpu() - ref).abs().max() <= 2e-5


@pytest.mark.parametrize("heads,hd", [(4, 32), (8, 16)])
def test_window_attention(be, heads, hd):
    """LocallyGroupedAttn core (twins.py:846-860) on a map whose size is not a multiple of the window: q, k, v live on the padded map."""
    lib, dev = be
    n, H, W, ws, C_ = 2, 10, 16, 7, heads * hd
    Hp, Wp = 14, 21
    q, k, v = rnd(n, Hp, Wp, C_, seed=1), rnd(n, Hp, Wp, C_, seed=2), rnd(n, Hp, Wp, C_, seed=3)
    out = torch.zeros(n, H, W, C_)
    qd, kd, vd, od = (t.to(dev) for t in (q, k, v, out))
    lib.check(lib.dll.gimmvfi_op_window_attention(C.byref(view_of(qd)), C.byref(view_of(kd)), C.byref(view_of(vd)), C.byref(view_of(od)), heads, ws, _stream(od)))
    sp = lambda t: t.reshape(n, Hp // ws, ws, Wp // ws, ws, heads, hd).permute(0, 1, 3, 5, 2, 4, 6).reshape(n, Hp // ws, Wp // ws, heads, ws * ws, hd)
    att = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) * hd ** -0.5, -1) @ sp(v)
    ref = att.reshape(n, Hp // ws, Wp // ws, heads, ws, ws, hd).permute(0, 1, 4, 2, 5, 3, 6).reshape(n, Hp, Wp, C_)[:, :H, :W]
    assert (od.cpu() - ref).abs().max() <= 1e-5


@pytest.mark.parametrize("heads,hd,hq,wq,hk,wk", [(4, 32, 48, 64, 6, 8), (8, 16, 64, 48, 7, 5), (8, 16, 8, 8, 2, 2)])
def test_global_attention(be, heads, hd, hq, wq, hk, wk):
    """GlobalSubSampleAttn core (twins.py:898-921).  >= 2048 queries take the shared-memory K/V kernel on the GPU (incl. a key count that
    is not a multiple of its group of 4); the small case the strided functor."""
    lib, dev = be
    n, C_ = 2, heads * hd
    q, kv = rnd(n, hq, wq, C_, seed=1), rnd(n, hk, wk, 2 * C_, seed=2)
    out = torch.zeros(n, hq, wq, C_)
    qd, kvd, od = q.to(dev), kv.to(dev), out.to(dev)
    lib.check(lib.dll.gimmvfi_op_global_attention(C.byref(view_of(qd)), C.byref(view_of(kvd, C_, 0)), C.byref(view_of(kvd, C_, C_)), C.byref(view_of(od)),
                                                  heads, _stream(od)))
    sh = lambda t: t.reshape(n, -1, heads, hd).transpose(1, 2)
    ref = (torch.softmax(sh(q) @ sh(kv[..., :C_]).transpose(-1, -2) * hd ** -0.5, -1) @ sh(kv[..., C_:])).transpose(1, 2).reshape(n, hq, wq, C_)
    assert (od.cpu() - ref).abs().max() <= 1e-5


@pytest.mark.parametrize("cin,k", [(3, 4), (128, 2), (64, 8)])
def test_patchify_is_strided_conv(be, cin, k):
    """patchify + a matmul with the re-laid-out weight == Conv2d(kernel = stride = k) (Twins PatchEmbed twins.py:1142-1149)."""
    lib, dev = be
    n, H, W, cout = 2, 16, 24, 8
    x, w = rnd(n, H, W, cin, seed=1), rnd(cout, cin, k, k, seed=2)
    ld = (cin + 3) // 4 * 4
    xp = torch.zeros(n, H, W, ld)
    xp[..., :cin] = x
    xd = xp.to(dev)
    od = torch.zeros(n, H // k, W // k, k * k * cin, device=dev)
    lib.check(lib.dll.gimmvfi_op_patchify(C.byref(view_of(xd, cin, 0)), C.byref(view_of(od)), k, _stream(od)))
    w2 = w.permute(0, 2, 3, 1).reshape(cout, k * k * cin)          # (ky*k + kx)*cin + ci
    got = od.cpu() @ w2.t()
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, stride=k).permute(0, 2, 3, 1)
    assert (got - ref).abs().max() <= 1e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("h,w", [(12, 16), (16, 20)])
def test_cost_conv1(be, h, w):
    """First layer of the cost-map patch embedding (encoder.py:38-48,68-71) incl. the zero extension to a multiple of 8 and the zero
    border the next x-packed stride-2 layer reads."""
    lib, dev = be
    maps = 5
    vol, wt, b = rnd(maps, h, w, seed=1), 0.2 * rnd(16, 1, 6, 6, seed=2), 0.1 * rnd(16, seed=3)
    Hp, Wp = (h + 7) // 8 * 8, (w + 7) // 8 * 8
    oh, ow = Hp // 2, Wp // 2
    od = torch.full((maps, oh + 4, ow + 4, 16), 3.0, device=dev)
    wt_t = wt.reshape(16, 36).t().contiguous()                      # [36 taps][16]
    vd = vol.to(dev)
    lib.check(lib.dll.gimmvfi_op_cost_conv1(P(vd), maps, h, w, P(wt_t), P(b), C.byref(view_of(od)), _stream(od)))
    ref = F.relu(F.conv2d(F.pad(vol[:, None], (0, Wp - w, 0, Hp - h)), wt, b, stride=2, padding=2)).permute(0, 2, 3, 1)
    got = od.cpu()
    assert (got[:, 2:-2, 2:-2] - ref).abs().max() <= 2e-5
    inner = torch.zeros_like(got, dtype=torch.bool)
    inner[:, 2:-2, 2:-2] = True
    assert got[~inner].abs().max() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,H,W", [(18, 3, 40, 70), (9, 4, 17, 33)])
def test_conv7x7_small_cout(cin, cout, H, W):
    """amt_comb_block.2 class (7x7, <= 4 output channels) on the shared-memory CUDA-core kernel: exact fp32 vs F.conv2d."""
    lib, dev = default_lib(), "cuda"
    n = 2
    x, w, b = rnd(n, H, W, cin, seed=1), 0.1 * rnd(cout, cin, 7, 7, seed=2), rnd(cout, seed=3)
    ld = (cin + 3) // 4 * 4
    xp = torch.zeros(n, H, W, ld)
    xp[..., :cin] = x
    pw = torch.zeros(49, cin, 4)
    pw[:, :, :cout] = w.permute(2, 3, 1, 0).reshape(49, cin, cout)
    bb = torch.zeros(4)
    bb[:cout] = b
    xd, pwd, bd = xp.to(dev), pw.to(dev), bb.to(dev)
    od = torch.zeros(n, H, W, 4, device=dev)
    lib.check(lib.dll.gimmvfi_op_conv7x7_small_cout(C.byref(view_of(xd, cin, 0)), P(pwd), P(bd), cout, C.byref(view_of(od, cout, 0)), _stream(od)))
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=3).permute(0, 2, 3, 1)
    assert (od.cpu()[..., :cout] - ref).abs().max() <= 2e-5
