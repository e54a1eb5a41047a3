"""CPU restatement of the index algebra of the Toeplitz-MMA snake kernel (flowhigh_b200/csrc/snake_mma.cuh) against the
oracle closed form: the banded up / down matrices with the kernel's k-slot -> time maps, the three-block scatter of the
down stage, and the error-feedback fp16 taps.  Pure numpy fp64 -- it pins the formulation, not the GPU arithmetic
(tests/test_gpu_kernels.py::test_snake_chunked does that)."""
import numpy as np
import torch

from oracle import model

F12 = np.array([0.00202896, 0.00938947, -0.02554346, -0.05765738, 0.12857258, 0.4432098,
                0.4432098, 0.12857258, -0.05765738, -0.02554346, 0.00938947, 0.00202896])


def tap(f, idx, scale):
    return scale * f[idx] if 0 <= idx < 12 else 0.0


def up_matrix(f, e, in16):
    """T_up[k-slot][n]: up-sample 2T + 8e + n from the 16 input steps of the K window (T - 8 + ko for e = 0, T + ko for
    e = 1).  fp32 input: slot 2q + 8r -> ko = q + 8r, slot 2q + 8r + 1 -> ko = q + 4 + 8r; fp16 (ldmatrix) input: identity."""
    base = 13 if e else 21
    T = np.zeros((16, 8))
    for q in range(4):
        for r in range(2):
            ko0 = 2 * q + 8 * r if in16 else q + 8 * r
            ko1 = ko0 + 1 if in16 else ko0 + 4
            for n in range(8):
                T[2 * q + 8 * r, n] = tap(f, n + base - 2 * ko0, 2.0)
                T[2 * q + 8 * r + 1, n] = tap(f, n + base - 2 * ko1, 2.0)
    return T


def slot_rows(in16):
    """time offset inside the 16-step K window held by k-slot c"""
    if in16:
        return np.arange(16)
    t8 = lambda c: (c >> 1) + 4 * (c & 1)
    return np.array([t8(c) if c < 8 else 8 + t8(c - 8) for c in range(16)])


def down_matrix(f, d):
    """T_dn,d[c][n]: output Q + n from up-sample 2Q + 16 (d - 1) + c"""
    T = np.zeros((16, 8))
    for c in range(16):
        for n in range(8):
            T[c, n] = tap(f, 16 * (d - 1) + c - 2 * n + 5, 1.0)
    return T


def snake_segment(xw, f, al, hib, NB, in16):
    """xw: the window of one half-segment, time q0 - 8 .. q0 + 8 NB + 8 (replicate-clamped by the caller); returns the
    8 NB outputs q0 .. q0 + 8 NB - 1, computed block by block exactly like the kernel (without the s~ edge fix-up)."""
    rows = slot_rows(in16)
    Tu = [up_matrix(f, 0, in16), up_matrix(f, 1, in16)]
    Td = [down_matrix(f, d) for d in range(3)]
    U = {}
    for j in range(-1, NB + 1):
        blk = np.zeros(16)
        for e in range(2):
            if (j == -1 and e == 0) or (j == NB and e == 1):
                continue  # only meets zero down-weights
            k0 = (8 * j if e else 8 * j - 8) + 8  # first window row of the K window
            u = xw[k0 + rows] @ Tu[e]
            blk[8 * e:8 * e + 8] = u - hib * np.cos(al * u)
        U[j] = blk
    y = np.full(8 * NB, hib)
    for j in range(-1, NB + 1):  # scatter: block j feeds down blocks j + 1 (first), j (middle), j - 1 (last term)
        for d, i in ((0, j + 1), (1, j), (2, j - 1)):
            if 0 <= i < NB:
                y[8 * i:8 * i + 8] += U[j] @ Td[d]
    return y


def test_toeplitz_blocks_equal_closed_form():
    rng = np.random.default_rng(0)
    L, NB = 700, 8
    x = rng.standard_normal(L) * 2
    alpha, beta = 0.2, -0.1
    filt = torch.from_numpy(F12).reshape(1, 1, 12)
    ref = model.aa_activation(torch.from_numpy(x).reshape(1, 1, L), torch.tensor([alpha]).double(),
                              torch.tensor([beta]).double(), filt, filt, True).numpy().reshape(L)
    al, hib = 2.0 * np.exp(alpha), 0.5 / (np.exp(beta) + 1e-9)
    xpad = np.concatenate([np.full(8, x[0]), x, np.full(8 * NB + 16, x[-1])])  # x~ for t = -8 ..
    for in16 in (False, True):
        for q0 in (0, 64, 320, 640):
            y = snake_segment(xpad[q0:q0 + 8 * NB + 16], F12, al, hib, NB, in16)
            lo, hi = max(q0, 3), min(q0 + 8 * NB, L - 3)  # outputs 0..2 / L-3..L-1 depend on the s~ clamp (scalar fix-up)
            # the sin^2 -> (1 - cos)/2 rewrite folds inv_b / 2 into the down filter assuming sum(f) = 1; the Kaiser taps
            # sum to 1 - 6e-8, hence the 4e-8 floor (same in the scalar kernel)
            assert np.abs(y[lo - q0:hi - q0] - ref[lo:hi]).max() < 2e-7, (in16, q0)


def test_slot_permutation_reads_contiguous_rows():
    """fp32 window: the four LDS.32 of a fragment read rows q, q + 4, q + 8, q + 12 -- per instruction the 32 lanes
    (g = channel, q = row) cover 4 consecutive 32-byte rows = 32 distinct banks."""
    rows = slot_rows(False)
    for s in range(4):  # register r = s >> 1, low / high half = s & 1
        slot = lambda q: 2 * q + 8 * (s >> 1) + (s & 1)
        words = sorted((rows[slot(q)] * 8 + g) % 32 for q in range(4) for g in range(8))
        assert words == list(range(32))


def ef_taps(f, scale):
    """error feedback inside each polyphase branch, largest tap first (snake_mma.cuh, CTA prologue)"""
    out = np.zeros(12)
    for ph in range(2):
        carry = 0.0
        for k in sorted(range(ph, 12, 2), key=lambda k: -abs(f[k])):
            v = np.float32(scale * f[k] + carry)
            h = np.float32(np.float16(v))
            carry, out[k] = float(v - h), float(h)
    return out


def test_error_feedback_taps_keep_branch_gains():
    f32 = F12.astype(np.float32).astype(np.float64)
    for scale in (2.0, 1.0):
        rn = np.float16(scale * f32).astype(np.float64)
        ef = ef_taps(f32, scale)
        assert np.all(np.float16(ef).astype(np.float64) == ef)  # representable in fp16
        for ph in range(2):
            exact = scale * f32[ph::2].sum()
            assert abs(ef[ph::2].sum() - exact) < 2e-6 < abs(rn[ph::2].sum() - exact)
        assert np.abs(ef - scale * f32).max() < 2.5e-4 * scale  # still within one fp16 ulp of the big taps
