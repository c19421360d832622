"""tcgen05 building blocks: fp16 (hi,lo)-split MMA with TMEM accumulators vs float64 matmul."""
import ctypes

import pytest
import torch

from mentflow_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [64, 128, 256])
def test_umma_split_gemm(n):
    lib = _lib.load()
    torch.manual_seed(n)
    a = (torch.randn(128, 64) * 3).cuda()
    b = (torch.randn(n, 64) * 0.2).cuda()
    a[:, 5] = 0.0
    d = torch.zeros(128, n, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.mfb_selftest_umma(a.data_ptr(), b.data_ptr(), n, d.data_ptr(), err.data_ptr(), st)
    assert rc == 0
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    want = a.double() @ b.double().T
    scale = (a.double().abs() @ b.double().abs().T)
    e = ((d.double() - want).abs() / scale).max()
    e32 = (((a @ b.T).double() - want).abs() / scale).max()
    assert float(e) < 2e-6, f"split-fp16 error {float(e):.2e} (fp32 matmul: {float(e32):.2e})"


@pytest.mark.parametrize("n,mode,kstep", [(64, 0, 0), (128, 0, 0), (64, 1, 0), (64, 1, 3), (128, 1, 2)])
def test_umma_sw32_tiles(n, mode, kstep):
    """One K step with 32-byte-row SWIZZLE_32B operand tiles (packed weight / bias tiles), also mixed
    with a SWIZZLE_128B A operand: exact for fp16-representable inputs."""
    lib = _lib.load()
    torch.manual_seed(n + mode)
    a = (torch.randint(-8, 9, (128, 16)).float() / 8).cuda()
    b = (torch.randint(-8, 9, (n, 16)).float() / 4).cuda()
    d = torch.zeros(128, n, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.mfb_selftest_umma_sw32(a.data_ptr(), b.data_ptr(), n, mode, kstep, d.data_ptr(), err.data_ptr(), st)
    assert rc == 0
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    assert torch.equal(d, a @ b.T)


@pytest.mark.parametrize("n", [64, 96, 128])
def test_umma_a_operand_in_tensor_memory(n):
    """the split GEMM with A written to TMEM by tcgen05.st and read from there by tcgen05.mma"""
    lib = _lib.load()
    torch.manual_seed(100 + n)
    a = (torch.randn(128, 64) * 3).cuda()
    b = (torch.randn(n, 64) * 0.2).cuda()
    d = torch.zeros(128, n, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.mfb_selftest_umma_ts(a.data_ptr(), b.data_ptr(), n, d.data_ptr(), err.data_ptr(), st)
    assert rc == 0
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    want = a.double() @ b.double().T
    scale = (a.double().abs() @ b.double().abs().T)
    e = ((d.double() - want).abs() / scale).max()
    assert float(e) < 2e-6, f"split-fp16 error {float(e):.2e}"
