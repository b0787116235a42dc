"""tcgen05 / TMA TF32 convolution (csrc/conv_tc.cu) vs PyTorch.

Strict check: against an fp64 convolution of the SAME operands the tensor core sees
(activations truncated to TF32 as the MMA does, weights rounded RN at pack time) — only
the fp32 accumulation order differs -> tolerance 5e-5 relative to the output scale.
Loose check: against the un-rounded fp32 convolution -> TF32 operand rounding, 3e-3."""
import pytest
import torch
import torch.nn.functional as F

import gpu_ops as K

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).to(DEV)


def act_fn(a, slope):
    return {0: lambda v: v, 1: F.relu, 2: lambda v: F.leaky_relu(v, 0.1), 3: lambda v: F.prelu(v, slope.double()), 6: torch.sin}[a]


CASES = [
    # cin, cout, k, H, W, n, act1
    (32, 16, 1, 8, 16, 1, 0),      # one tile, one K block, 1x1: the minimal MMA
    (32, 16, 1, 16, 32, 2, 0),     # several tiles / persistent loop
    (128, 64, 1, 16, 32, 1, 1),    # 4 K blocks
    (64, 64, 3, 16, 32, 1, 0),     # 3x3 halo via TMA OOB zero fill
    (64, 64, 3, 20, 28, 2, 2),     # ragged tiles (H, W not multiples of 8 / 16)
    (256, 256, 3, 24, 40, 1, 3),   # the dominant final-decoder shape, N=256, PReLU
    (273, 256, 3, 16, 24, 1, 3),   # cin tail block zero-filled
    (64, 64, 5, 16, 32, 1, 3),     # 5x5
    (256, 24, 3, 16, 32, 1, 0),    # cout 24 -> N=32 padded
    (256, 188, 3, 16, 20, 1, 2),   # cout 188 -> N=192
    (648, 256, 1, 16, 20, 2, 2),   # K=648 1x1 (corr features)
    (35, 128, 1, 24, 24, 1, 6),    # HypoNet layer 0, sin
    (8, 32, 5, 32, 32, 1, 3),      # tiny cin
]


@pytest.mark.parametrize("case", CASES + [(256, 576, 1, 16, 20, 1, 0)])
def test_conv2d_tc(case):
    cin, cout, k, H, W, n, act1 = case
    x = rnd(n, cin, H, W, seed=1)
    w = rnd(cout, cin, k, k, seed=2, scale=1.0 / (cin * k * k) ** 0.5)
    b = rnd(cout, seed=3, scale=0.1)
    slope = (0.25 + 0.1 * rnd(cout, seed=4)) if act1 == 3 else None
    xn = K.nhwc(x)
    if cin % 4:  # TMA needs 16-byte pixel strides: the engine pads such buffers (ld 276 / 36), so does the test
        pad = torch.full((n, H, W, (cin + 3) // 4 * 4), 7.0, device=DEV)   # non-zero padding lanes must be ignored
        pad[..., :cin] = xn
        got = K.nchw(K.conv2d_tc(pad, w, b, act1, slope, in_view=K.view_of(pad, channels=cin)))
    else:
        got = K.nchw(K.conv2d_tc(xn, w, b, act1, slope))
    f = act_fn(act1, slope)
    strict = f(F.conv2d(K.tf32_trunc(x).double(), K.tf32_rn(w).double(), b.double(), padding=k // 2)).float()
    loose = f(F.conv2d(x.double(), w.double(), b.double(), padding=k // 2)).float()
    e_strict = (got - strict).abs().max().item()
    e_loose = (got - loose).abs().max().item()
    print("case", case, "strict %.3e loose %.3e ref absmax %.3e" % (e_strict, e_loose, loose.abs().max().item()))
    # outputs are stored rounded to TF32 (RN): |err| <= 2^-11 * |value| on top of the accumulation-order term
    assert e_strict <= 5e-5 + 2.0 ** -11 * loose.abs().max().item()
    assert e_loose <= 3e-3 + 2.0 ** -10 * loose.abs().max().item()


def test_conv2d_tc_two_segments_residual_act2():
    """ResBlock conv5: prelu(x + conv(cat[b[:, :192], s2]))  (fi_components.py:147-153)."""
    a = rnd(1, 256, 16, 32, seed=1)
    s = rnd(1, 64, 16, 32, seed=2)
    w = rnd(256, 256, 3, 3, seed=3, scale=0.02)
    b = rnd(256, seed=4, scale=0.1)
    res = rnd(1, 256, 16, 32, seed=5)
    slope = 0.25 + 0.1 * rnd(256, seed=6)
    a_nhwc = K.nhwc(a)
    got = K.nchw(K.conv2d_tc(a_nhwc, w, b, 0, None, K.nhwc(res), 3, slope, x1_nhwc=K.nhwc(s), in_view=K.view_of(a_nhwc, channels=192)))
    xin = torch.cat([a[:, :192], s], 1)
    ref = F.prelu(res.double() + F.conv2d(K.tf32_trunc(xin).double(), K.tf32_rn(w).double(), b.double(), padding=1), slope.double()).float()
    assert (got - ref).abs().max().item() <= 5e-5 + 2.0 ** -11 * ref.abs().max().item()


def test_conv2d_tc_channel_slice_views():
    """Input = channel slice of a wider buffer (ld > c), output into a slice."""
    buf = rnd(1, 276, 16, 32, seed=1)
    w = rnd(64, 64, 3, 3, seed=2, scale=0.05)
    b = rnd(64, seed=3, scale=0.1)
    buf_nhwc = K.nhwc(buf)
    got = K.nchw(K.conv2d_tc(buf_nhwc, w, b, in_view=K.view_of(buf_nhwc, channels=64, offset=128)))
    ref = F.conv2d(K.tf32_trunc(buf[:, 128:192]).double(), K.tf32_rn(w).double(), b.double(), padding=1).float()
    assert (got - ref).abs().max().item() <= 5e-5 + 2.0 ** -11 * ref.abs().max().item()


SPLIT_CASES = [
    # cin, cout, kh, kw, H, W, n, act1
    (64, 64, 3, 3, 16, 32, 1, 0),
    (384, 128, 1, 5, 16, 20, 2, 4),     # SepConvGRU horizontal gate (sigmoid)
    (384, 128, 5, 1, 16, 20, 2, 5),     # vertical, tanh
    (324, 256, 1, 1, 16, 20, 2, 1),     # lookup features -> 256
    (256, 126, 3, 3, 16, 20, 1, 1),     # cout 126
    (2, 128, 7, 7, 16, 20, 2, 1),       # flow encoder: cin 2 (ld 4)
    (256, 576, 1, 1, 16, 20, 1, 0),     # mask head: cout 576 -> 3 N tiles of 192
    (256, 2, 3, 3, 16, 20, 2, 0),       # flow head: cout 2
    (128, 256, 3, 3, 24, 40, 1, 1),
    (128, 133, 3, 3, 24, 40, 1, 0),     # 2 N tiles whose width is not a multiple of the 32-column epilogue chunk (init-decoder head)
    (96, 96, 3, 3, 24, 40, 2, 1),       # cin not a multiple of the 64-element K block of the 3xF16 form
    (272, 128, 1, 1, 24, 40, 1, 0),
]


@pytest.mark.parametrize("f16", [False, True], ids=["3xtf32", "3xf16"])
@pytest.mark.parametrize("case", SPLIT_CASES)
def test_conv2d_tc_3xtf32(case, f16):
    """Three-term operand splitting (3xTF32, and 3xF16: fp16 hi / lo pairs on kind::f16): fp32-class accuracy (vs an fp64
    convolution of the UNROUNDED operands)."""
    cin, cout, kh, kw, H, W, n, act1 = case
    x = rnd(n, cin, H, W, seed=1)
    w = rnd(cout, cin, kh, kw, seed=2, scale=1.0 / (cin * kh * kw) ** 0.5)
    b = rnd(cout, seed=3, scale=0.1)
    xn = K.nhwc(x)
    if cin % 4:
        pad = torch.full((n, H, W, (cin + 3) // 4 * 4), 7.0, device=DEV)
        pad[..., :cin] = xn
        got = K.nchw(K.conv2d_tc(pad, w, b, act1, in_view=K.view_of(pad, channels=cin), split=True, split_f16=f16))
    else:
        got = K.nchw(K.conv2d_tc(xn, w, b, act1, split=True, split_f16=f16))
    f = {0: lambda v: v, 1: F.relu, 4: torch.sigmoid, 5: torch.tanh}[act1]
    ref = f(F.conv2d(x.double(), w.double(), b.double(), padding=(kh // 2, kw // 2))).float()
    err = (got - ref).abs().max().item()
    print("split case", case, "3xf16" if f16 else "3xtf32", "err %.3e ref absmax %.3e" % (err, ref.abs().max().item()))
    # operands are exact to ~2^-22; what remains is the tensor core's fp32 accumulation (long chains, truncating adder)
    assert err <= 1e-4


@pytest.mark.parametrize("f16", [False, True], ids=["3xtf32", "3xf16"])
def test_conv2d_tc_gru_epilogues(f16):
    """r-gate: sigmoid(conv) * h;  q-gate: h' = (1-z) h + z tanh(conv(cat[r*h, x]))  (raft/update.py:52-59)."""
    n, H, W = 2, 16, 20
    hx = rnd(n, 384, H, W, seed=1)
    wr = rnd(128, 384, 1, 5, seed=2, scale=0.03)
    wq = rnd(128, 384, 1, 5, seed=3, scale=0.03)
    br, bq = rnd(128, seed=4, scale=0.1), rnd(128, seed=5, scale=0.1)
    z = torch.sigmoid(rnd(n, 128, H, W, seed=6))
    hxn = K.nhwc(hx)
    h_view = K.view_of(hxn, channels=128)
    # r * h
    hbuf = K.nhwc(hx[:, :128])
    rh = K.conv2d_tc(hxn, wr, br, 4, mul=hbuf, split=True, split_f16=f16)
    rh_ref = torch.sigmoid(F.conv2d(hx.double(), wr.double(), br.double(), padding=(0, 2))).float() * hx[:, :128]
    assert (K.nchw(rh) - rh_ref).abs().max().item() <= 1e-4
    # q with two input segments and the GRU blend
    xin = K.nhwc(hx[:, 128:])
    got = K.conv2d_tc(rh, wq, bq, 5, x1_nhwc=xin, gru_z=K.nhwc(z), gru_h=hbuf, split=True, split_f16=f16)
    q = torch.tanh(F.conv2d(torch.cat([rh_ref, hx[:, 128:]], 1).double(), wq.double(), bq.double(), padding=(0, 2))).float()
    ref = (1 - z) * hx[:, :128] + z * q
    assert (K.nchw(got) - ref).abs().max().item() <= 1e-4


def test_conv2d_tc_gru_hoisted_epilogues_cluster():
    """SepConvGRU with the context term hoisted (engine.cu: `_hm` / `_inp` weights) at a size that runs on CTA pairs:
    two input segments [h | motion], a pre-activation residual, sigmoid / tanh as act2, gate multiply on the r half
    (merged z | r output) and the GRU blend  (raft/update.py:52-66)."""
    n, H, W = 2, 136, 240
    h = rnd(n, 128, H, W, seed=1); mot = rnd(n, 128, H, W, seed=2)
    wq = rnd(128, 256, 1, 5, seed=3, scale=0.03); bq = torch.zeros(128, device=DEV)
    pq = rnd(n, 128, H, W, seed=4, scale=0.5)
    z = torch.sigmoid(rnd(n, 128, H, W, seed=6))
    rh = rnd(n, 128, H, W, seed=7)
    hn = K.nhwc(h)
    got = K.conv2d_tc(K.nhwc(rh), wq, bq, 0, residual=K.nhwc(pq), act2=5, x1_nhwc=K.nhwc(mot), gru_z=K.nhwc(z), gru_h=hn, split=True, split_f16=True)
    q = torch.tanh(F.conv2d(torch.cat([rh, mot], 1).double(), wq.double(), None, padding=(0, 2)) + pq.double()).float()
    ref = (1 - z) * h + z * q
    err = (K.nchw(got) - ref).abs().max().item()
    print("hoisted q gate err %.3e" % err)
    assert err <= 1e-4
    # vertical half, plain sigmoid gate with the residual
    wz = rnd(128, 256, 5, 1, seed=8, scale=0.03)
    got = K.conv2d_tc(hn, wz, bq, 0, residual=K.nhwc(pq), act2=4, x1_nhwc=K.nhwc(mot), split=True, split_f16=True)
    ref = torch.sigmoid(F.conv2d(torch.cat([h, mot], 1).double(), wz.double(), None, padding=(2, 0)) + pq.double()).float()
    err = (K.nchw(got) - ref).abs().max().item()
    print("hoisted z gate err %.3e" % err)
    assert err <= 1e-4


CLUSTER_CASES = [
    # >= 2 pixel tiles per SM -> the 2-CTA weight-multicast path; 168x240 = 21x15 = 315 tiles (odd: exercises padding)
    (64, 64, 3, 3, 168, 240, 1, 0, False),
    (256, 256, 3, 3, 168, 240, 1, 3, False),
    (96, 192, 3, 3, 160, 256, 1, 1, False),
    (384, 128, 1, 5, 168, 240, 1, 4, True),
    (128, 256, 3, 3, 168, 240, 1, 1, True),     # split, 2 N tiles: both CTAs of a cluster must share the N tile
    (64, 64, 3, 3, 160, 256, 2, 0, True),
    (256, 128, 5, 1, 168, 240, 1, 5, True),     # vertical gate, tanh, on CTA pairs
    (96, 96, 3, 3, 170, 250, 1, 1, True),       # ragged tiles + a K block whose upper half is beyond the tensor (TMA zero fill)
    (28, 64, 7, 1, 168, 240, 1, 1, True),       # 7 taps (the x-packed 7x7 stem's shape class)
]


@pytest.mark.parametrize("f16", [False, True], ids=["3xtf32", "3xf16"])
@pytest.mark.parametrize("case", CLUSTER_CASES)
def test_conv2d_tc_cluster_multicast(case, f16):
    cin, cout, kh, kw, H, W, n, act1, split = case
    if f16 and not split:
        pytest.skip("3xF16 is a form of the split kernel")
    x = rnd(n, cin, H, W, seed=1)
    w = rnd(cout, cin, kh, kw, seed=2, scale=1.0 / (cin * kh * kw) ** 0.5)
    b = rnd(cout, seed=3, scale=0.1)
    slope = (0.25 + 0.1 * rnd(cout, seed=4)) if act1 == 3 else None
    got = K.nchw(K.conv2d_tc(K.nhwc(x), w, b, act1, slope, split=split, split_f16=f16))
    f = {0: lambda v: v, 1: F.relu, 3: lambda v: F.prelu(v, slope.double()), 4: torch.sigmoid, 5: torch.tanh}[act1]
    if split:
        ref = f(F.conv2d(x.double(), w.double(), b.double(), padding=(kh // 2, kw // 2))).float()
        tol = 1e-4
    else:
        ref = f(F.conv2d(K.tf32_trunc(x).double(), K.tf32_rn(w).double(), b.double(), padding=(kh // 2, kw // 2))).float()
        tol = 5e-5 + 2.0 ** -11 * ref.abs().max().item()
    err = (got - ref).abs().max().item()
    print("cluster case", case, "err %.3e" % err)
    assert err <= tol


F16_CASES = [
    # cin, cout, k, H, W, n
    (64, 64, 1, 8, 16, 1),        # one tile, one 64-channel K block
    (256, 256, 3, 24, 40, 1),     # the residual trunk's dominant shape (N = 256)
    (64, 64, 3, 20, 28, 2),       # side conv, ragged tiles
    (256, 24, 3, 16, 32, 1),      # trunk -> 24 output channels (fp32 out in the network)
]


@pytest.mark.parametrize("case", F16_CASES)
@pytest.mark.parametrize("out_half", [True, False])
def test_conv2d_tc_f16_operands(case, out_half):
    """kind::f16 path (precision mode 3): half activations and weights, fp32 accumulation.  Strict reference = fp64
    convolution of the SAME half-rounded operands; the only differences are accumulation order and the output rounding."""
    cin, cout, k, H, W, n = case
    x = rnd(n, cin, H, W, seed=1)
    w = rnd(cout, cin, k, k, seed=2, scale=1.0 / (cin * k * k) ** 0.5)
    b = rnd(cout, seed=3, scale=0.1)
    slope = 0.25 + 0.1 * rnd(cout, seed=4)
    xh = K.nhwc(x).half()
    got = K.conv2d_tc_f16(xh, w, b, 3, slope, out_half=out_half)
    assert got.dtype == (torch.float16 if out_half else torch.float32)
    ref = F.prelu(F.conv2d(x.half().double(), w.half().double(), b.double(), padding=k // 2), slope.double()).float()
    err = (K.nchw(got.float()) - ref).abs().max().item()
    tol = 1e-3 * max(1.0, ref.abs().max().item())   # half store, or fp32 store rounded to TF32 for the next layer: 2^-11 relative
    print("f16 case", case, out_half, "err %.3e" % err, "ref absmax %.3e" % ref.abs().max().item())
    assert err <= tol


def test_conv2d_tc_f16_resblock_tail():
    """two half input segments (192 + 64 channels), half residual, PReLU after the add, half output: conv5 of
    ResBlock(256, 64) (fi_components.py:139-154) as the engine runs it in precision mode 3"""
    n, H, W = 1, 24, 40
    a = rnd(n, 192, H, W, seed=1); s2 = rnd(n, 64, H, W, seed=2); res = rnd(n, 256, H, W, seed=3)
    w = rnd(256, 256, 3, 3, seed=4, scale=1.0 / (256 * 9) ** 0.5)
    b = rnd(256, seed=5, scale=0.1)
    slope = 0.25 + 0.1 * rnd(256, seed=6)
    ah, sh, rh = K.nhwc(a).half(), K.nhwc(s2).half(), K.nhwc(res).half()
    got = K.conv2d_tc_f16(ah, w, b, 0, None, residual=rh, act2=3, slope2=slope, x1_nhwc=sh, out_half=True)
    xin = torch.cat([a.half().double(), s2.half().double()], 1)
    ref = F.prelu(F.conv2d(xin, w.half().double(), b.double(), padding=1) + res.half().double(), slope.double()).float()
    err = (K.nchw(got.float()) - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item())


def test_conv2d_tc_tf32_in_half_out():
    """fp32 / TF32 operands with a half-precision store: convblock.0 (273 -> 256) feeding the half trunk"""
    n, cin, cout, H, W = 1, 273, 256, 16, 24
    x = rnd(n, cin, H, W, seed=1)
    w = rnd(cout, cin, 3, 3, seed=2, scale=1.0 / (cin * 9) ** 0.5)
    b = rnd(cout, seed=3, scale=0.1)
    slope = 0.25 + 0.1 * rnd(cout, seed=4)
    pad = torch.zeros(n, H, W, 276, device=DEV); pad[..., :cin] = K.nhwc(x)
    lib = K.default_lib()
    import ctypes as C
    pw, pwh = K.pack_weight_tc(w), K.pack_weight_tc_f16(w)
    bb = torch.zeros(512, device=DEV); bb[:cout] = b
    out = torch.empty(n, H, W, cout, device=DEV, dtype=torch.float16)
    P = lambda t: C.c_void_p(t.data_ptr())
    lib.check(lib.dll.gimmvfi_op_conv2d_tc_f16(C.byref(K.view_of(pad, channels=cin)), None, P(pwh), P(pw), P(bb), cin, cout, 3, 3, 3, P(slope), None, 0,
                                               None, 2, C.byref(K.view_of(out)), K._stream(out)))
    ref = F.prelu(F.conv2d(K.tf32_trunc(x).double(), K.tf32_rn(w).double(), b.double(), padding=1), slope.double()).float()
    err = (K.nchw(out.float()) - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [(64, 96, 3, 32, 48, 2), (64, 96, 1, 32, 48, 1), (96, 128, 3, 40, 56, 2), (64, 64, 3, 34, 62, 1)])
@pytest.mark.parametrize("split", [False, True])
def test_conv2d_tc_stride2(case, split):
    """stride-2 convolutions of the RAFT encoder (raft/extractor.py:42-48): TMA element strides gather every 2nd pixel"""
    import ctypes as C
    cin, cout, k, H, W, n = case
    x = rnd(n, cin, H, W, seed=1)
    w = rnd(cout, cin, k, k, seed=2, scale=1.0 / (cin * k * k) ** 0.5)
    b = rnd(cout, seed=3, scale=0.1)
    xn = K.nhwc(x)
    Ho, Wo = (H + 2 * (k // 2) - k) // 2 + 1, (W + 2 * (k // 2) - k) // 2 + 1
    out = torch.empty(n, Ho, Wo, cout, device=DEV)
    pw = K.pack_weight_tc(w)
    bb = torch.zeros(512, device=DEV); bb[:cout] = b
    lib = K.default_lib()
    lib.check(lib.dll.gimmvfi_op_conv2d_tc_strided(C.byref(K.view_of(xn)), C.c_void_p(pw.data_ptr()), C.c_void_p(bb.data_ptr()), cin, cout, k, k, 2, 1,
                                                   int(split), C.byref(K.view_of(out)), K._stream(out)))
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=k // 2)).float()
    assert ref.shape[2:] == (Ho, Wo)
    err = (K.nchw(out) - ref).abs().max().item()
    tol = (2e-5 if split else 3e-3) * max(1.0, ref.abs().max().item())
    print("stride-2 case", case, "split", split, "err %.3e" % err)
    assert err <= tol


@pytest.mark.parametrize("hw", [(16, 24), (72, 96)])   # 72x96: 54 source tiles x 54 target tiles -> CTA pairs (>= 2 x 148 tiles)
@pytest.mark.parametrize("split", [True, False])
def test_corr_volume_tc(hw, split):
    """all-pairs correlation (raft/corr.py:167-175) on the tcgen05 kernel: 3xTF32 for RAFT, TF32 for the bidirectional volume"""
    import ctypes as C
    h, w = hw
    Cc = 256
    g = torch.Generator().manual_seed(3)
    fa = torch.randn(1, h, w, Cc, generator=g).to(DEV)
    fb = torch.randn(1, h, w, Cc, generator=g).to(DEV)
    N = h * w
    vol = torch.empty(N, N, device=DEV)
    scratch = torch.empty(2 * N * Cc + N + 2048, device=DEV)
    lib = K.default_lib()
    lib.check(lib.dll.gimmvfi_op_corr_volume_tc(C.byref(K.view_of(fa)), C.byref(K.view_of(fb)), C.c_void_p(scratch.data_ptr()), C.c_void_p(vol.data_ptr()),
                                                int(split), K._stream(vol)))
    a = fa.view(N, Cc).double(); b = fb.view(N, Cc).double()
    if not split:
        a, b = K.tf32_trunc(fa).view(N, Cc).double(), K.tf32_trunc(fb).view(N, Cc).double()
    ref = (a @ b.t() / 16.0).float()
    err = (vol - ref).abs().max().item()
    print("corr_volume_tc", hw, "split", split, "err %.3e" % err, "absmax %.2f" % ref.abs().max().item())
    assert err <= (3e-5 if split else 2e-4)


HALO_CASES = [
    # cin, cout, H, W, n, act1, residual, half
    (32, 32, 48, 40, 2, 2, False, False),     # the 32 -> 32 full-resolution layers (LeakyReLU)
    (32, 32, 50, 37, 1, 0, True, False),      # ragged tiles (8 x 16), residual epilogue
    (16, 32, 32, 24, 1, 2, False, False),     # cin 16: K block zero-filled by TMA
    (32, 64, 32, 24, 2, 3, False, False),     # BN = 64, PReLU
    (32, 16, 34, 26, 1, 0, False, False),     # cout 16
    (64, 64, 32, 24, 1, 3, False, True),      # half-precision storage: 64 channels = one K block
    (64, 64, 35, 20, 1, 0, True, True),
    (64, 64, 40, 56, 2, 1, False, True, 1),   # 1x1 (final-decoder upsample.7): one tap, the box is the 8 x 16 tile itself
    (32, 48, 33, 20, 1, 3, False, False, 1),  # 1x1 fp32 / TF32, ragged tiles
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv2d_halo(case):
    """csrc/conv_halo.cu: the tile + halo is loaded once and every tap is an MMA on a shifted view (UMMA base offset).  Strict
    reference = fp64 convolution of the operands as the tensor core sees them (TF32-truncated activations / RN weights, or halves)."""
    cin, cout, H, W, n, act1, use_res, half = case[:8]
    ks = case[8] if len(case) > 8 else 3
    x = rnd(n, cin, H, W, seed=1)
    w = rnd(cout, cin, ks, ks, seed=2, scale=1.0 / (cin * ks * ks) ** 0.5)
    b = rnd(cout, seed=3, scale=0.1)
    slope = (0.25 + 0.1 * rnd(cout, seed=4)) if act1 == 3 else None
    res = rnd(n, cout, H, W, seed=5) if use_res else None
    xn = K.nhwc(x)
    if half:
        got = K.conv2d_halo(xn.half(), w, b, act1, slope, residual=K.nhwc(res).half() if use_res else None, out_half=True)
        ref = F.conv2d(x.half().double(), w.half().double(), b.double(), padding=ks // 2)
    else:
        got = K.conv2d_halo(xn, w, b, act1, slope, residual=K.nhwc(res) if use_res else None)
        ref = F.conv2d(K.tf32_trunc(x).double(), K.tf32_rn(w).double(), b.double(), padding=ks // 2)
    f = {0: lambda v: v, 1: F.relu, 2: lambda v: F.leaky_relu(v, 0.1), 3: lambda v: F.prelu(v, slope.double())}[act1]
    ref = f(ref)
    if use_res:
        ref = ref + (res.half().double() if half else res.double())
    ref = ref.float()
    err = (K.nchw(got.float()) - ref).abs().max().item()
    tol = (1e-3 if half else 5e-5 + 2.0 ** -11) * max(1.0, ref.abs().max().item())   # store rounding: half / RN to TF32
    print("halo case", case, "err %.3e (tol %.1e)" % (err, tol))
    assert err <= tol


def test_conv2d_halo_prepadded_matches_reflect():
    """reflect-padded layers (cnn_encoder.7): valid convolution on a pre-padded buffer, tile origin without the -1 shift"""
    x = rnd(1, 32, 40, 32, seed=1)
    w = rnd(16, 32, 3, 3, seed=2, scale=0.06)
    b = rnd(16, seed=3, scale=0.1)
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    got = K.nchw(K.conv2d_halo(K.nhwc(xp), w, b, prepadded=True))
    ref = F.conv2d(K.tf32_trunc(xp).double(), K.tf32_rn(w).double(), b.double()).float()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 5e-5 + 2.0 ** -11 * ref.abs().max().item()
