"""CPU emulation of fh_tc_conv's addressing (tile decode, chunk windows, packed weight image,
tap-shifted descriptors, epilogue indexing) against the plain tapped convolution.  This pins the
host-side packing (flowhigh_b200/packing.py) and the index arithmetic of csrc/tc_conv.cu; the
tensor-core semantics themselves are checked on the GPU (tests/test_gpu_kernels.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from flowhigh_b200 import packing

HALO = 32


def emulate_tc_conv(a_chunked, a_bs, a_cs, a_row0, packed, rec_off, P, ntaps, cin_pad, cout_pad, bn, B, L,
                    out, o_bs, o_cs, o_row, bias=None):
    """Follows tc_conv_kernel step by step (fp32 arithmetic on the bf16-rounded operands)."""
    m_tiles, n_tiles, ci_pairs = -(-L // 128), -(-cout_pad // bn), -(-cin_pad // 16)
    n_chunks = cin_pad // 8  # odd chunk count: the partner window of the last chunk is a zeroed smem region (ci_odd)
    a = a_chunked.float().numpy()
    w = packed.float().numpy().reshape(P, n_tiles, ci_pairs, ntaps, 2, bn, 8)
    for b in range(B):
        for p in range(P):
            offs = rec_off[p]
            mn = int(offs.min())
            wrows = 128 + int(offs.max() - offs.min())
            for mt in range(m_tiles):
                for nt in range(n_tiles):
                    D = np.zeros((128, bn), np.float32)
                    row = a_row0 + mt * 128 + mn
                    for cp in range(ci_pairs):
                        slot = np.zeros((2, wrows, 8), np.float32)
                        for c in range(2):
                            if 2 * cp + c >= n_chunks:
                                continue
                            base = b * a_bs + (2 * cp + c) * a_cs + row * 8
                            slot[c] = a[base: base + wrows * 8].reshape(wrows, 8)
                        for j in range(ntaps):
                            sh = int(offs[j]) - mn
                            A = np.concatenate([slot[0, sh: sh + 128], slot[1, sh: sh + 128]], axis=1)  # [128,16]
                            Bm = np.concatenate([w[p, nt, cp, j, 0], w[p, nt, cp, j, 1]], axis=1)      # [bn,16]
                            D += A @ Bm.T
                    for r in range(128):
                        t = mt * 128 + r
                        if t >= L:
                            continue
                        for n in range(bn):
                            ng = nt * bn + n
                            if ng >= cout_pad:
                                continue
                            v = D[r, n] + (0.0 if bias is None else float(bias[ng]))
                            out[b * o_bs + (ng // 8) * o_cs + (t * P + p) * o_row + ng % 8] = v


def to_chunked(x, cpad, Lp, row0):
    """x [B,C,L] -> flat chunked bf16 buffer [B][cpad/8][Lp][8]."""
    B, C, L = x.shape
    buf = torch.zeros(B, cpad // 8, Lp, 8)
    xp = torch.zeros(B, cpad, L)
    xp[:, :C] = x
    buf[:, :, row0: row0 + L, :] = xp.reshape(B, cpad // 8, 8, L).permute(0, 1, 3, 2)
    return buf.to(torch.bfloat16).flatten()


def from_chunked(flat, B, C, cpad, Lp, row0, L):
    t = torch.from_numpy(flat).reshape(B, cpad // 8, Lp, 8)[:, :, row0: row0 + L, :]
    return t.permute(0, 1, 3, 2).reshape(B, cpad, L)[:, :C]


@pytest.mark.parametrize("kind", ["conv_d1", "conv_d5", "convT_5_11", "convT_2_4", "linear"])
def test_tc_addressing_matches_conv(kind):
    torch.manual_seed(0)
    B, Cin, Cout, L = 2, 24, 40, 150
    x = torch.randn(B, Cin, L)
    bias = torch.randn(Cout)
    if kind.startswith("conv_d"):
        d = int(kind[-1])
        w = torch.randn(Cout, Cin, 7) * 0.2
        tc = packing.conv1d_taps(w, bias, d)
        ref = F.conv1d(x.bfloat16().float(), w.bfloat16().float(), bias, dilation=d, padding=(7 * d - d) // 2)
    elif kind.startswith("convT"):
        _, u, k = kind.split("_")
        u, k = int(u), int(k)
        w = torch.randn(Cin, Cout, k) * 0.2
        tc = packing.conv_transpose1d_taps(w, bias, u)
        ref = F.conv_transpose1d(x.bfloat16().float(), w.bfloat16().float(), bias, stride=u, padding=(k - u) // 2)
    else:
        w = torch.randn(Cout, Cin) * 0.2
        tc = packing.linear_taps(w, bias)
        ref = F.conv1d(x.bfloat16().float(), w.bfloat16().float()[:, :, None], bias)
    packed, cin_pad, cout_pad, bn = packing.pack_tc(tc, "cpu", bn=32)
    assert cin_pad == 24 and cout_pad == 40  # whole 8-channel chunks; the weight image pads Cin to whole ci-pairs
    assert packed.numel() * 2 == tc.P * (-(-cout_pad // bn)) * (-(-cin_pad // 16)) * tc.ntaps * bn * 32
    Lp_in = HALO + packing.round_up(L, 128) + 64
    Lo = L * tc.P
    Lp_out = HALO + packing.round_up(Lo, 128) + 64
    a = to_chunked(x, cin_pad, Lp_in, HALO)
    a = torch.cat([a, torch.zeros(4096, dtype=a.dtype)])
    out = np.zeros(B * (cout_pad // 8) * Lp_out * 8, np.float32)
    emulate_tc_conv(a, (cin_pad // 8) * Lp_in * 8, Lp_in * 8, HALO, packed, tc.off, tc.P, tc.ntaps, cin_pad, cout_pad, bn,
                    B, L, out[HALO * 8:], (cout_pad // 8) * Lp_out * 8, Lp_out * 8, 8,
                    bias=packing.pad_vec(bias, cout_pad).numpy())
    got = from_chunked(out, B, Cout, cout_pad, Lp_out, HALO, Lo)
    assert torch.allclose(got, ref, atol=2e-3, rtol=1e-3), (got - ref).abs().max()
    # halo rows were never written
    full = torch.from_numpy(out).reshape(B, cout_pad // 8, Lp_out, 8)
    assert full[:, :, :HALO].abs().max() == 0 and full[:, :, HALO + Lo:].abs().max() == 0


def _mma(D, a_img, a_start, a_lbo, b_img, b_start, b_lbo, bn):
    """One 128 x bn x 16 MMA from flat shared-memory images of 8-element (16-byte) rows, addressed like the no-swizzle
    K-major descriptors of tc_conv.cu: K half h of A row r = row (a_start + h * a_lbo + r), of B column n = row
    (b_start + h * b_lbo + n); all offsets in 16-byte rows."""
    for h in range(2):
        A = a_img[a_start + h * a_lbo: a_start + h * a_lbo + 128]      # [128, 8]
        Bm = b_img[b_start + h * b_lbo: b_start + h * b_lbo + bn]      # [bn, 8]
        D += A @ Bm.T


@pytest.mark.parametrize("k,d", [(7, 1), (11, 1), (3, 5), (4, 3), (2, 1)])
def test_tc_tap_pair_addressing(k, d):
    """Odd chunk count (24 channels): mma_role issues the taps of the odd chunk two per MMA -- A: LBO = one tap step on the
    same window, B: LBO = one tap slot of the UNCHANGED weight image -- and an odd last tap in the old form (zero partner
    window).  Emulated on flat shared-memory images against the plain convolution."""
    torch.manual_seed(k * 10 + d)
    Cin, Cout, L, bn = 24, 24, 128, 32
    x = torch.randn(1, Cin, L)
    w = torch.randn(Cout, Cin, k) * 0.2
    tc = packing.conv1d_taps(w, None, d)
    ref = F.conv1d(x.bfloat16().float(), w.bfloat16().float(), None, dilation=d, padding=(k * d - d) // 2)
    packed, cin_pad, cout_pad, _ = packing.pack_tc(tc, "cpu", bn=bn)
    offs = tc.off[0]
    mn, span = int(offs.min()), int(offs.max() - offs.min())
    step = int(offs[1] - offs[0])
    assert step == d and all(int(offs[j + 1] - offs[j]) == step for j in range(k - 1))  # arithmetic taps (tap_arith)
    wrows = 128 + span
    arows_pad = 128 + 64
    Lp = HALO + 128 + 64
    a = to_chunked(x, cin_pad, Lp, HALO).float().numpy().reshape(cin_pad // 8, Lp, 8)
    wimg = packed.float().numpy().reshape(2, k, 2, bn, 8)  # [ci-pair][tap][half][bn][8] (P = 1, one N tile)
    D = np.zeros((128, bn), np.float32)
    rel0 = int(offs[0]) - mn
    for cp in range(2):
        # stage image: two chunk windows of arows_pad rows (the partner of the odd chunk is a zeroed region), then the weights
        slot = np.zeros((2 * arows_pad, 8), np.float32)
        for c in range(2):
            if 2 * cp + c < cin_pad // 8:
                slot[c * arows_pad: c * arows_pad + wrows] = a[2 * cp + c, HALO + mn: HALO + mn + wrows]
        bimg = wimg[cp].reshape(k * 2 * bn, 8)
        if cp == 0:  # full pair: one MMA per tap, K halves = the two chunks (LBO = chunk window stride / half a tap slot)
            for j in range(k):
                _mma(D, slot, rel0 + j * step, arows_pad, bimg, j * 2 * bn, bn, bn)
        else:        # odd chunk: two taps per MMA
            for j in range(k // 2):
                _mma(D, slot, rel0 + 2 * j * step, step, bimg, 2 * j * 2 * bn, 2 * bn, bn)
            if k & 1:
                _mma(D, slot, rel0 + (k - 1) * step, arows_pad, bimg, (k - 1) * 2 * bn, bn, bn)
    got = torch.from_numpy(D[:, :Cout].T.copy())[None][..., : ref.shape[-1]]  # an even k yields one output less in torch
    assert torch.allclose(got, ref, atol=2e-3, rtol=1e-3), (got - ref).abs().max()
