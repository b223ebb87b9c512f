"""Parity of the tcgen05 implicit-GEMM conv kernel (through the C ABI) against a torch fp32 reference.

Tolerance: outputs are stored in bf16 (8-bit mantissa), accumulation is fp32 on both sides, so
|got - ref| <= 2^-7 |ref| + 2e-2 covers one bf16 rounding plus summation-order noise.
"""
import pytest
import torch

from gpu_util import conv_reference, describe_mismatch

pytestmark = pytest.mark.gpu

RTOL, ATOL = 2.0 ** -7, 2e-2

# (B, H, W, in_ctot, in_c_off, cin, cout, k, stride, act, residual, out_ctot, out_c_off, out_fp32)
CASES = [
    (1, 16, 16, 64, 0, 64, 64, 1, 1, 1, False, 64, 0, False),
    (2, 20, 20, 128, 0, 128, 128, 1, 1, 1, False, 128, 0, False),
    (1, 16, 24, 48, 0, 48, 80, 1, 1, 0, False, 80, 0, True),
    (1, 8, 8, 32, 16, 16, 16, 1, 1, 1, False, 48, 32, False),
    (1, 16, 16, 64, 0, 64, 64, 3, 1, 1, False, 64, 0, False),
    (2, 20, 20, 96, 32, 64, 32, 3, 1, 1, True, 96, 64, False),
    (1, 32, 32, 32, 0, 32, 64, 3, 2, 1, False, 64, 0, False),
    (2, 40, 40, 192, 64, 128, 128, 3, 2, 1, False, 160, 16, False),
    (1, 20, 20, 256, 0, 256, 512, 3, 1, 1, False, 512, 0, False),
    (1, 23, 40, 80, 0, 80, 176, 3, 1, 1, False, 176, 0, False),
    (4, 80, 80, 128, 0, 128, 128, 3, 1, 1, True, 128, 0, False),
    (3, 46, 80, 160, 0, 160, 320, 3, 2, 1, False, 320, 0, False),
    (2, 40, 40, 384, 0, 384, 128, 1, 1, 1, False, 128, 0, False),
    (8, 160, 160, 32, 0, 32, 64, 3, 1, 1, False, 64, 0, False),     # > 148 tiles: several tiles per persistent CTA
    (16, 80, 80, 64, 0, 64, 320, 1, 1, 1, False, 320, 0, False),    # Cout split across tiles, TMEM double buffering
    (6, 40, 40, 128, 0, 128, 256, 3, 1, 1, True, 256, 0, False),    # n_tile 256: both accumulators fill TMEM
    (2, 32, 32, 64, 0, 64, 16, 1, 1, 1, False, 16, 0, False),       # n_tile 16: one epilogue warp per lane group idles
    (3, 20, 20, 64, 0, 64, 80, 1, 1, 0, False, 176, 64, True),      # fp32 head rows: 80 classes into a 176-float row
    (2, 24, 24, 48, 0, 48, 144, 3, 1, 1, True, 144, 0, False),      # 9 chunks: uneven column split, residual
    (16, 80, 80, 96, 0, 96, 64, 1, 1, 1, False, 64, 0, False),      # 800 flat tiles: 256-pixel CTA tiles (two sub-tiles)
    (13, 80, 80, 64, 0, 64, 128, 1, 1, 1, True, 192, 64, False),    # same, odd tile count, residual, output slice
    (8, 160, 160, 32, 0, 32, 64, 3, 2, 1, False, 64, 0, False),     # stride 2 with two stacked sub-tiles per CTA tile
    (20, 80, 80, 64, 0, 64, 128, 3, 2, 1, False, 128, 0, False),    # stride 2, 40x40 output, 120-row sub-tiles
    (3, 46, 74, 32, 0, 32, 64, 3, 2, 1, False, 96, 32, False),      # stride-2 halo (Cin 32): ragged 23x37 output, output slice
    (2, 64, 64, 32, 0, 32, 32, 3, 2, 1, False, 32, 0, False),       # stride-2 halo, n_tile 32
    (5, 32, 48, 32, 0, 32, 48, 3, 2, 0, False, 48, 0, True),        # stride-2 halo, no activation, fp32 rows
]


def run_case(case, impl=0, seed=0):
    from yolo_puncture_b200.engine import conv2d_bf16, gemm_weight
    B, H, W, ictot, icoff, cin, cout, k, s, act, use_res, octot, ocoff, ofp32 = case
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    x = torch.randn((B, H, W, ictot), device=dev, generator=g).to(torch.bfloat16)
    w = torch.randn((cout, cin, k, k), device=dev, generator=g) * (1.0 / (cin * k * k) ** 0.5)
    bias = torch.randn((cout,), device=dev, generator=g) * 0.5
    oH, oW = H // s, W // s
    odt = torch.float32 if ofp32 else torch.bfloat16
    out = torch.full((B, oH, oW, octot), 7.0, device=dev, dtype=odt)
    res = None
    if use_res:
        out = torch.randn((B, oH, oW, octot), device=dev, generator=g).to(odt)
        res = out.clone()
    conv2d_bf16(x, gemm_weight(w), bias, k, s, act, cin=cin, in_c_off=icoff, res=res, out=out, out_c_off=ocoff,
                out_fp32=ofp32, impl=impl)
    torch.cuda.synchronize()
    ref = conv_reference(x[..., icoff:icoff + cin], w, bias, k, s, act,
                         res=None if res is None else res[..., ocoff:ocoff + cout])
    got = out[..., ocoff:ocoff + cout].float()
    untouched = torch.cat([out[..., :ocoff], out[..., ocoff + cout:]], -1)
    untouched_ref = torch.cat([res[..., :ocoff], res[..., ocoff + cout:]], -1) if use_res else torch.full_like(untouched, 7.0)
    return got, ref, bool((untouched == untouched_ref).all())


@pytest.mark.parametrize("impl", [0, 2], ids=["persistent", "one_tile_per_cta"])
@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_conv_tc_matches_reference(case, impl):
    got, ref, clean = run_case(case, impl=impl)
    ok = bool(((got - ref).abs() <= ATOL + RTOL * ref.abs()).all())
    assert ok, describe_mismatch(got, ref, RTOL, ATOL)
    assert clean, "kernel wrote outside its channel slice"


@pytest.mark.parametrize("case", CASES[:8], ids=[str(i) for i in range(8)])
def test_conv_simt_twin_matches_reference(case):
    got, ref, clean = run_case(case, impl=1)
    ok = bool(((got - ref).abs() <= ATOL + RTOL * ref.abs()).all())
    assert ok, describe_mismatch(got, ref, RTOL, ATOL)
    assert clean


# CTA-pair kernel (conv3_halo2_kernel: tcgen05.mma.cta_group::2, weights shared by the two SMs of a TPC).  The planner
# picks it for 3x3 stride-1 layers with streamed weights and >= 74 tile pairs; YPB_PAIR_MIN=1 forces it on small cases.
PAIR_CASES = [
    (8, 80, 80, 128, 0, 128, 128, 3, 1, 1, False, 128, 0, False),      # many pairs per cluster, K = 2 full chunks
    (4, 80, 80, 320, 160, 160, 160, 3, 1, 1, True, 320, 160, False),   # ragged third chunk (4, 4, 2 K-steps), residual, slices
    (3, 40, 40, 320, 0, 320, 320, 3, 1, 1, False, 320, 0, False),      # two Cout splits of 160; 45 tiles: the last pair is half empty
    (2, 23, 40, 96, 16, 80, 80, 3, 1, 1, True, 80, 0, False),          # K-steps (4, 1), odd map size
    (5, 20, 20, 256, 0, 256, 224, 3, 1, 0, False, 224, 0, True),       # n_tile 224, fp32 rows, no activation
    (2, 160, 160, 128, 0, 128, 128, 3, 1, 1, False, 128, 0, False),    # 400 tiles per image
]


@pytest.mark.parametrize("case", PAIR_CASES, ids=[str(i) for i in range(len(PAIR_CASES))])
def test_conv_pair_kernel_matches_reference(case, monkeypatch):
    monkeypatch.setenv("YPB_PAIR_MIN", "1")
    got, ref, clean = run_case(case, impl=0, seed=3)
    ok = bool(((got - ref).abs() <= ATOL + RTOL * ref.abs()).all())
    assert ok, describe_mismatch(got, ref, RTOL, ATOL)
    assert clean, "kernel wrote outside its channel slice"
    monkeypatch.setenv("YPB_NO_PAIR", "1")  # and the single-CTA halo kernel on the same case, bit for bit the same result
    got1, _, _ = run_case(case, impl=0, seed=3)
    assert torch.equal(got, got1)


def test_planner_picks_the_pair_kernel_for_wide_layers():
    import ctypes as C
    from yolo_puncture_b200._lib import check, diag_lib
    B, H, W, cin, cout = 32, 80, 80, 160, 160
    x = torch.zeros((B, H, W, cin), device="cuda", dtype=torch.bfloat16)
    w = torch.zeros((9, cout, cin), device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros((cout,), device="cuda")
    out = torch.empty((B, H, W, cout), device="cuda", dtype=torch.bfloat16)
    ms, desc = C.c_float(), C.create_string_buffer(640)
    check(diag_lib().ypb_conv_bench(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(x.data_ptr()), B, H, W, cin, 0, cin,
                                    C.c_void_p(w.data_ptr()), C.c_void_p(bias.data_ptr()), cout, 3, 1, 1, None,
                                    C.c_void_p(out.data_ptr()), cout, 0, 0, 0, -1, 2, C.byref(ms), desc, 640))
    assert desc.value.decode().startswith("pair(cta_group::2)"), desc.value.decode()


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_every_case_with_cta_pairs_forced(case, monkeypatch):
    """All shapes of the main table again with YPB_PAIR_MIN=1: 1x1 and stride-2 3x3 layers then run on conv_tc2p_kernel, 3x3
    stride-1 layers with streamed weights on conv3_halo2_kernel (resident-weight halo layers keep their kernel); results
    must equal the single-CTA kernels' bit for bit (same accumulation order) and match the fp32 reference."""
    monkeypatch.setenv("YPB_PAIR_MIN", "1")
    got, ref, clean = run_case(case, impl=0)
    ok = bool(((got - ref).abs() <= ATOL + RTOL * ref.abs()).all())
    assert ok, describe_mismatch(got, ref, RTOL, ATOL)
    assert clean, "kernel wrote outside its channel slice"
    monkeypatch.setenv("YPB_NO_PAIR", "1")
    got1, _, _ = run_case(case, impl=0)
    assert torch.equal(got, got1)
