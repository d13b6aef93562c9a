"""GPU (-m gpu): hardware self-test of the tcgen05 descriptor patterns used by the bf16 denoiser."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def sw128_off(r, c):
    return (r >> 3) * 1024 + (r & 7) * 128 + (((c ^ r) & 7) << 4)


def make_image(mat_bf16):
    """mat: [rows, 64] bf16 (as torch) -> 128B-swizzled byte image (rows multiple of 8)."""
    rows = mat_bf16.shape[0]
    raw = mat_bf16.view(torch.int16).numpy().reshape(rows, 8, 8)           # [row, chunk, 8 elems]
    img = np.zeros(rows * 128, dtype=np.uint8)
    for r in range(rows):
        for c in range(8):
            off = sw128_off(r, c)
            img[off:off + 16] = raw[r, c].view(np.uint8)
    return img


def run_selftest(a_img, b_img, a_start_off, sbo, N, nk16, base_offset=0):
    from cld_b200._lib import lib
    a = torch.from_numpy(a_img).cuda()
    b = torch.from_numpy(b_img).cuda()
    d = torch.zeros(128, N, device="cuda")
    rc = lib.cld_tc_selftest(C.c_void_p(a.data_ptr()), a.numel(), C.c_void_p(b.data_ptr()), b.numel(), a_start_off,
                             sbo, N, nk16, base_offset, C.c_void_p(d.data_ptr()),
                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.cld_last_error(None)
    torch.cuda.synchronize()
    return d.cpu()


@pytest.mark.parametrize("N,nk16,shift_slots,sbo", [(64, 4, 0, 1024), (128, 4, 2, 1024), (256, 4, 4, 1024),
                                                    (64, 1, 1, 1024), (128, 4, 1, 2048), (256, 2, 0, 2048),
                                                    (16, 4, 3, 1024)])
def test_umma_shifted_and_strided_descriptors(N, nk16, shift_slots, sbo):
    torch.manual_seed(N + nk16 + shift_slots)
    a_rows = 8 * 40
    A = torch.randn(a_rows, 64).bfloat16()
    B = torch.randn(N, 64).bfloat16()
    d = run_selftest(make_image(A), make_image(B), shift_slots * 1024, sbo, N, nk16)
    K = 16 * nk16
    m = torch.arange(128)
    rows = shift_slots * 8 + (m // 8) * (sbo // 1024) * 8 + (m % 8)
    want = A[rows, :K].float() @ B[:, :K].float().t()
    err = (d - want).abs().max().item()
    assert err < 1e-3 * max(1.0, want.abs().max().item()), err


def test_umma_row_shift_inside_swizzle_atom_report():
    """Informational: does a 128-byte (one row) start offset work, with or without base_offset?
    The denoiser does not rely on it (it shifts by whole 1024-byte atoms); the result is printed."""
    torch.manual_seed(5)
    A = torch.randn(8 * 40, 64).bfloat16()
    B = torch.randn(64, 64).bfloat16()
    m = torch.arange(128)
    for shift_rows in (1, 3):
        want = A[shift_rows + m, :64].float() @ B.float().t()
        for bo in (0, shift_rows):
            d = run_selftest(make_image(A), make_image(B), shift_rows * 128, 1024, 64, 4, base_offset=bo)
            ok = (d - want).abs().max().item() < 1e-3 * want.abs().max().item()
            print("row shift %d, base_offset %d: %s" % (shift_rows, bo, "matches" if ok else "differs"))
