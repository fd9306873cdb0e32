"""The tcgen05 / TMEM / TMA GEMM core in isolation (b200f_umma_selftest) against torch fp32 matmul on the
same 16-bit operands, for every operand layout the head uses, plus K1's fp16 operand output."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(dev, M, N, K, a_mn, b_mn, fmt=2, k_splits=1, pair=2):
    import b200face
    from b200face import _lib
    lib = b200face.load_library()
    old_pair = lib.b200f_umma_set_pair(pair)
    try:
        return _run_pair(lib, _lib, dev, M, N, K, a_mn, b_mn, fmt, k_splits)
    finally:
        lib.b200f_umma_set_pair(old_pair)


def _run_pair(lib, _lib, dev, M, N, K, a_mn, b_mn, fmt, k_splits):
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g, device=dev)
    b = torch.randn(N, K, generator=g, device=dev)
    a16 = a.half() if fmt >= 1 else a.bfloat16()
    b16 = b.half() if fmt == 2 else b.bfloat16()
    ref = a16.float() @ b16.float().t()
    a_s = a16.t().contiguous() if a_mn else a16.contiguous()
    b_s = b16.t().contiguous() if b_mn else b16.contiguous()
    out = torch.full((k_splits, M, N), float("nan"), device=dev)
    _lib.check(lib.b200f_umma_selftest(_lib.ptr(a_s), _lib.ptr(b_s), _lib.ptr(out), M, N, K, a_mn, b_mn, fmt, k_splits,
                                       -1, -1, -1, -1, -1, -1, _lib.stream_ptr(dev)), "umma_selftest")
    torch.cuda.synchronize()
    assert lib.b200f_umma_timeout_flag(1) == 0, "a bounded pipeline wait expired"
    return float((out.sum(0) - ref).norm() / ref.norm())


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (512, 1024, 512), (296, 704, 192), (8, 8, 8), (1000, 264, 72)])
@pytest.mark.parametrize("fmt", [0, 2])
@pytest.mark.parametrize("pair", [1, 2])
def test_gemm_core_layouts(cuda_device, a_mn, b_mn, M, N, K, fmt, pair):
    """pair = 1: cta_group::1 on 128 x 256 tiles; pair = 2: cta_group::2 CTA pairs on 256 x 256 tiles."""
    assert _run(cuda_device, M, N, K, a_mn, b_mn, fmt, pair=pair) < 2e-6


@pytest.mark.parametrize("pair", [1, 2])
def test_gemm_core_split_k(cuda_device, pair):
    assert _run(cuda_device, 512, 512, 8192, 0, 1, 2, k_splits=16, pair=pair) < 2e-6
    assert _run(cuda_device, 512, 512, 20000, 1, 1, 2, k_splits=18, pair=pair) < 2e-6     # the dX GEMM's layouts


@pytest.mark.parametrize("pair", [1, 2])
def test_gemm_core_many_tiles_persistent(cuda_device, pair):
    """More work items than SMs: the persistent loop, both accumulator stages and the smem ring wrap."""
    assert _run(cuda_device, 1024, 148 * 256 + 512, 256, 0, 0, 2, pair=pair) < 2e-6


@pytest.mark.parametrize("pair", [1, 2])
@pytest.mark.parametrize("B,C,D", [(128, 128, 64), (256, 256, 512), (512, 4096, 512), (300, 1000, 200), (8, 8, 8),
                                   (1024, 148 * 300, 512), (130, 70000, 512)])
def test_xw_kernel(cuda_device, B, C, D, pair):
    """The X-stationary kernel that carries K2 / K3a (x_hat resident in shared memory, class tiles streamed), on
    single CTAs and on tcgen05 cta_group::2 CTA pairs: raw accumulators against torch fp32 matmul."""
    import b200face
    from b200face import _lib
    lib = b200face.load_library()
    g = torch.Generator(device=cuda_device).manual_seed(B + C + D)
    x = torch.randn(B, D, generator=g, device=cuda_device).half()
    w = torch.randn(C, D, generator=g, device=cuda_device).half()
    out = torch.full((B, C), float("nan"), device=cuda_device)
    _lib.check(lib.b200f_umma_xw_selftest(_lib.ptr(x), _lib.ptr(w), _lib.ptr(out), B, C, D, pair,
                                          _lib.stream_ptr(cuda_device)), "umma_xw_selftest")
    torch.cuda.synchronize()
    assert lib.b200f_umma_timeout_flag(1) == 0, "a bounded pipeline wait expired"
    ref = x.float() @ w.float().t()
    assert float((out - ref).norm() / ref.norm()) < 2e-6


@pytest.mark.parametrize("in_dt", [torch.float32, torch.bfloat16])
def test_k1_emits_normalised_fp16_operands(cuda_device, in_dt):
    from b200face.head import OPERAND_SCALE, _k1
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(333, 512, generator=g) * 5).to(in_dt)
    xo, inv = _k1(x.to(cuda_device).contiguous(), True)
    assert xo.dtype == torch.float16
    ref = torch.nn.functional.normalize(x.float().double(), dim=1) * OPERAND_SCALE
    np.testing.assert_allclose(xo.float().cpu().double().numpy(), ref.numpy(), rtol=2 ** -11, atol=2 ** -24)
    np.testing.assert_allclose(inv.cpu().double().numpy(), (1 / x.float().double().norm(dim=1)).numpy(), rtol=3e-7)
